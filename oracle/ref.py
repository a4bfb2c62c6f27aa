"""ctypes front end of the REFERENCE's own object code (oracle/_ref/libref_*.so, built by
oracle/build_ref.sh from /root/reference).  TEST INFRASTRUCTURE.

The reference functions are K&R C over fixed-size arrays; the strides below are the (patched)
capacity #defines of each build variant (oracle/_ref/variants.txt):
  alpha/beta/symbol_probab : [MAX_STATES][MAX_TIME]         (T-FS:112-114)
  transition_probab        : [MAX_STATES][MAX_STATES]
  gaus_probab_dens (1 frame): [MAX_STATES][MAX_MIXTURE]
  struct state             : { double mix_coef[MAX_MIX];
                               struct { double mean[MAX_COEF], cov[MAX_COEF], det; } mix[MAX_MIX]; }
MAX_PARAMETERS_NUMBER is 1 in every variant, so symbol_probab[P][N][T] == [N][T].
"""
import ctypes as C
import os
import struct
import subprocess

import numpy as np

from .oracle import Model, Stats

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
MAX_TIME = 500
MAX_STATES_TRAIN = 20
MAX_STATES_TEST = 15
VARIANTS = {"stock": (9, 3), "d39m16": (39, 16), "d39m128": (39, 128)}  # tag -> (MAX_COEF, MAX_MIX)

_dp = C.POINTER(C.c_double)


def available(tag="stock"):
    return os.path.exists(os.path.join(REF_DIR, "libref_train_%s.so" % tag))


def _d(a):
    return a.ctypes.data_as(_dp)


class RefTrain:
    """Function-level access to the trainer T-FS (hmm_continuous_fs.c)."""

    def __init__(self, tag="stock"):
        self.tag = tag
        self.MC, self.MM = VARIANTS[tag]
        self.MS = MAX_STATES_TRAIN
        self.lib = C.CDLL(os.path.join(REF_DIR, "libref_train_%s.so" % tag))
        for f in ("calc_gaus", "calc_probability", "calc_det", "classifying"):
            getattr(self.lib, f).restype = C.c_double
        self.state_doubles = self.MM + self.MM * (2 * self.MC + 1)

    # ---- struct state packing ----
    def pack_states(self, m):
        buf = np.zeros((self.MS, self.state_doubles))
        for i in range(m.N):
            buf[i, : m.M] = m.c[i]
            for j in range(m.M):
                o = self.MM + j * (2 * self.MC + 1)
                buf[i, o : o + m.D] = m.mu[i, j]
                buf[i, o + self.MC : o + self.MC + m.D] = m.iv[i, j]
                buf[i, o + 2 * self.MC] = m.det[i, j]
        return buf

    def unpack_states(self, buf, N, M, D):
        c = np.zeros((N, M)); mu = np.zeros((N, M, D)); iv = np.zeros((N, M, D)); det = np.zeros((N, M))
        for i in range(N):
            c[i] = buf[i, :M]
            for j in range(M):
                o = self.MM + j * (2 * self.MC + 1)
                mu[i, j] = buf[i, o : o + D]
                iv[i, j] = buf[i, o + self.MC : o + self.MC + D]
                det[i, j] = buf[i, o + 2 * self.MC]
        return c, mu, iv, det

    def pack_A(self, A):
        buf = np.zeros((self.MS, self.MS))
        buf[: A.shape[0], : A.shape[1]] = A
        return buf

    # ---- A2 ----
    def calc_gaus(self, x, mu, iv, det):
        x = np.ascontiguousarray(x, dtype=np.float64)
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        iv = np.ascontiguousarray(iv, dtype=np.float64)
        return self.lib.calc_gaus(C.c_int(len(x)), _d(x), _d(mu), _d(iv), C.c_double(det))

    # ---- A3: all frames -> b[T][N], post[T][N][M] ----
    def emissions(self, m, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        T = x.shape[0]
        assert T <= MAX_TIME
        st = self.pack_states(m)
        sp = np.zeros((self.MS, MAX_TIME))
        post = np.zeros((T, m.N, m.M))
        gbuf = np.zeros((self.MS, self.MM))
        for t in range(T):
            xt = np.ascontiguousarray(x[t])
            self.lib.calc_symbol_probab(C.c_int(m.N), C.c_int(m.M), C.c_int(m.D), _d(xt), _d(st),
                                        _d(gbuf), _d(sp), C.c_int(t))
            post[t] = gbuf[: m.N, : m.M]
        return sp[: m.N, :T].T.copy(), post, sp

    # ---- A4..A9 over one utterance, accumulating into `acc` (dict of reference-shaped arrays) ----
    def new_acc(self):
        return dict(num=np.zeros((self.MS, self.MS)), den=np.zeros(self.MS), denmix=np.zeros(self.MS),
                    nmp=np.zeros((self.MS, self.state_doubles)))

    def utterance(self, m, x, acc):
        """Runs calc_symbol_probab/alpha/beta/transition/den_mix/mix_param/probability exactly in the
        order of the trainer's main loop (T-FS:272-319).  Returns dict(b, post, alpha, beta, scale, logp)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        T = x.shape[0]
        b, post, sp = self.emissions(m, x)
        st = self.pack_states(m)
        A = self.pack_A(m.A)
        alpha = np.zeros((self.MS, MAX_TIME)); beta = np.zeros((self.MS, MAX_TIME)); sc = np.zeros(MAX_TIME)
        pi = np.zeros(self.MS, dtype=np.int32); pi[0] = 1
        L = self.lib
        L.calc_alpha(C.c_int(m.N), C.c_int(T), C.c_int(1), _d(alpha), _d(sc), _d(A), _d(sp),
                     pi.ctypes.data_as(C.POINTER(C.c_int)))
        L.calc_beta(C.c_int(m.N), C.c_int(T), C.c_int(1), _d(beta), _d(sc), _d(A), _d(sp))
        L.calc_transition_probab(C.c_int(m.N), C.c_int(T), C.c_int(1), _d(alpha), _d(beta), _d(sc), _d(A),
                                 _d(sp), _d(acc["num"]), _d(acc["den"]))
        L.calc_den_mix_coef(C.c_int(T), C.c_int(m.N), _d(alpha), _d(beta), _d(sc), _d(acc["denmix"]))
        gbuf = np.zeros((self.MS, self.MM))
        for t in range(T):
            gbuf[:] = 0
            gbuf[: m.N, : m.M] = post[t]
            xt = np.ascontiguousarray(x[t])
            L.calc_mix_param(C.c_int(t), C.c_int(m.N), C.c_int(m.M), C.c_int(m.D), _d(xt), _d(alpha),
                             _d(beta), _d(sc), _d(gbuf), _d(acc["nmp"]), _d(st))
        lp = L.calc_probability(C.c_int(T), _d(sc), C.c_double(alpha[m.N - 1, T - 1]))
        return dict(b=b, post=post, alpha=alpha[: m.N, :T].T.copy(), beta=beta[: m.N, :T].T.copy(),
                    scale=sc[:T].copy(), logp=lp)

    def acc_to_stats(self, acc, N, M, D):
        s = Stats(N, M, D)
        s.num_trans = acc["num"][:N, :N].copy()
        s.den_trans = acc["den"][:N].copy()
        s.den_mix = acc["denmix"][:N].copy()
        s.S0, s.S1, s.S2c, _ = self.unpack_states(acc["nmp"], N, M, D)
        return s

    # ---- A10 (the M-step block of main(), T-FS:328-346) ----
    def mstep(self, m, acc):
        st = self.pack_states(m)
        A = self.pack_A(m.A)
        L = self.lib
        L.updating_transition_probab(C.c_int(m.N), _d(acc["num"]), _d(acc["den"]), _d(A))
        L.updating_mix_param(C.c_int(m.N), C.c_int(m.M), C.c_int(m.D), _d(acc["denmix"]), _d(acc["nmp"]), _d(st))
        for i in range(m.N):
            for j in range(m.M):
                o = self.MM + j * (2 * self.MC + 1)
                cov = st[i, o + self.MC : o + 2 * self.MC]  # view
                st[i, o + 2 * self.MC] = L.calc_det(C.c_int(m.D), _d(cov))
                L.inv_matrix(C.c_int(m.D), _d(cov))
        c, mu, iv, det = self.unpack_states(st, m.N, m.M, m.D)
        return Model(A[: m.N, : m.N], c, mu, iv, det, m.word)

    # ---- initial-model builder through the reference's own entry point (T-FS:732) ----
    def init_model(self, N, M, list_file, word="w"):
        data_file = C.create_string_buffer(100)
        data_file.value = list_file.encode()
        mixn = (C.c_int * 1)(M)
        coefn = (C.c_int * 1)(0)
        A = np.zeros((self.MS, self.MS)); st = np.zeros((self.MS, self.state_doubles))
        self.lib.creating_initial_model(C.c_int(1), data_file, C.c_int(N), mixn, coefn, _d(A), _d(st))
        D = coefn[0]
        c, mu, iv, det = self.unpack_states(st, N, M, D)
        return Model(A[:N, :N], c, mu, iv, det, word)


class RefTest:
    """Function-level access to the recogniser R-FS (recognition_continuous_fs.c)."""

    def __init__(self, tag="stock"):
        self.MC, self.MM = VARIANTS[tag]
        self.MS = MAX_STATES_TEST
        self.lib = C.CDLL(os.path.join(REF_DIR, "libref_test_%s.so" % tag))
        self.lib.calc_probability.restype = C.c_double
        self.lib.calc_gaus.restype = C.c_double
        self.state_doubles = self.MM + self.MM * (2 * self.MC + 1)

    def forward_score(self, m, x):
        """R-FS:349-367 for one (utterance, model) cell."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        T = x.shape[0]
        st = np.zeros((self.MS, self.state_doubles))
        for i in range(m.N):
            st[i, : m.M] = m.c[i]
            for j in range(m.M):
                o = self.MM + j * (2 * self.MC + 1)
                st[i, o : o + m.D] = m.mu[i, j]
                st[i, o + self.MC : o + self.MC + m.D] = m.iv[i, j]
                st[i, o + 2 * self.MC] = m.det[i, j]
        A = np.zeros((self.MS, self.MS)); A[: m.N, : m.N] = m.A
        sp = np.zeros((self.MS, MAX_TIME)); alpha = np.zeros((self.MS, MAX_TIME)); sc = np.zeros(MAX_TIME)
        pi = np.zeros(self.MS, dtype=np.int32); pi[0] = 1
        for t in range(T):
            xt = np.ascontiguousarray(x[t])
            self.lib.calc_symbol_probab(C.c_int(m.N), C.c_int(m.M), C.c_int(m.D), _d(xt), _d(st), _d(sp), C.c_int(t))
        self.lib.calc_alpha(C.c_int(m.N), C.c_int(T), C.c_int(1), _d(alpha), _d(sc), _d(A), _d(sp),
                            pi.ctypes.data_as(C.POINTER(C.c_int)))
        return self.lib.calc_probability(C.c_int(T), _d(sc), C.c_double(alpha[m.N - 1, T - 1]))

    def rank(self, score):
        score = np.ascontiguousarray(score, dtype=np.float64)
        idx = np.zeros(len(score), dtype=np.int32)
        self.lib.sorting_probab(_d(score), idx.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(len(score)))
        return idx


# ---------------- the reference's on-disk formats (numpy side, for fixtures) ----------------
def write_features(path, x):
    """Feature file: int32 D, then T x D doubles (T-FS:527-581)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    with open(path, "wb") as f:
        f.write(struct.pack("<i", x.shape[1]))
        f.write(x.tobytes())


def read_features(path):
    with open(path, "rb") as f:
        raw = f.read()
    D = struct.unpack("<i", raw[:4])[0]
    n = (len(raw) - 4) // (8 * D)
    return np.frombuffer(raw, dtype="<f8", count=n * D, offset=4).reshape(n, D).copy()


def read_model(path, len_bytes=8):
    """.hmm layout (T-FS:2058-2144).  len_bytes=4 reads the shipped 32-bit files' header."""
    with open(path, "rb") as f:
        raw = f.read()
    p = 0
    ln = struct.unpack("<Q" if len_bytes == 8 else "<I", raw[p : p + len_bytes])[0]; p += len_bytes
    word = raw[p : p + ln].decode(); p += ln
    N, P = struct.unpack("<ii", raw[p : p + 8]); p += 8
    assert P == 1
    M = struct.unpack("<i", raw[p : p + 4])[0]; p += 4
    D = struct.unpack("<i", raw[p : p + 4])[0]; p += 4
    def take(n):
        nonlocal p
        a = np.frombuffer(raw, dtype="<f8", count=n, offset=p).copy(); p += 8 * n
        return a
    A = take(N * N).reshape(N, N)
    c = np.zeros((N, M)); mu = np.zeros((N, M, D)); iv = np.zeros((N, M, D)); det = np.zeros((N, M))
    for i in range(N):
        c[i] = take(M)
        for j in range(M):
            mu[i, j] = take(D); det[i, j] = take(1)[0]; iv[i, j] = take(D)
    assert p == len(raw), (p, len(raw))
    return Model(A, c, mu, iv, det, word)


def read_model_streams(path, len_bytes=8):
    """.hmm with P >= 1 feature streams -> [Model per stream] (all carrying the same A and word)."""
    with open(path, "rb") as f:
        raw = f.read()
    p = 0
    ln = struct.unpack("<Q" if len_bytes == 8 else "<I", raw[p : p + len_bytes])[0]; p += len_bytes
    word = raw[p : p + ln].decode(); p += ln
    N, P = struct.unpack("<ii", raw[p : p + 8]); p += 8
    Ms = struct.unpack("<%di" % P, raw[p : p + 4 * P]); p += 4 * P
    Ds = struct.unpack("<%di" % P, raw[p : p + 4 * P]); p += 4 * P
    def take(n):
        nonlocal p
        a = np.frombuffer(raw, dtype="<f8", count=n, offset=p).copy(); p += 8 * n
        return a
    A = take(N * N).reshape(N, N)
    out = []
    for M, D in zip(Ms, Ds):
        c = np.zeros((N, M)); mu = np.zeros((N, M, D)); iv = np.zeros((N, M, D)); det = np.zeros((N, M))
        for i in range(N):
            c[i] = take(M)
            for j in range(M):
                mu[i, j] = take(D); det[i, j] = take(1)[0]; iv[i, j] = take(D)
        out.append(Model(A, c, mu, iv, det, word))
    assert p == len(raw), (p, len(raw))
    return out


def write_model(path, m):
    with open(path, "wb") as f:
        w = m.word.encode()
        f.write(struct.pack("<Q", len(w))); f.write(w)
        f.write(struct.pack("<iiii", m.N, 1, m.M, m.D))
        f.write(np.ascontiguousarray(m.A).tobytes())
        for i in range(m.N):
            f.write(np.ascontiguousarray(m.c[i]).tobytes())
            for j in range(m.M):
                f.write(np.ascontiguousarray(m.mu[i, j]).tobytes())
                f.write(struct.pack("<d", m.det[i, j]))
                f.write(np.ascontiguousarray(m.iv[i, j]).tobytes())


def run_train_cli(tag, word, N, M, list_file, out_hmm, cwd=None, stack_unlimited=False):
    exe = os.path.join(REF_DIR, "hmm_fs_%s" % tag)
    cmd = "%s %s %d 1 %d %s %s" % (exe, word, N, M, list_file, out_hmm)
    if stack_unlimited:
        cmd = "ulimit -s unlimited; " + cmd
    subprocess.run(["bash", "-c", cmd], cwd=cwd, stdout=subprocess.DEVNULL, check=True)


def run_test_cli(tag, models_list, feat_list, words_file, result_file, cwd=None, capture=False):
    exe = os.path.join(REF_DIR, "rec_fs_%s" % tag)
    r = subprocess.run([exe, "1", models_list, "1", feat_list, words_file, result_file], cwd=cwd,
                       stdout=subprocess.PIPE if capture else subprocess.DEVNULL, check=True)
    return r.stdout.decode(errors="replace") if capture else None


def parse_train_report(txt_path):
    mean = its = None
    for line in open(txt_path):
        if line.startswith("mean probability:"):
            mean = float(line.split(":")[1])
        if line.startswith("number of iterations:"):
            its = int(line.split(":")[1])
    return mean, its


def run_train_cli_timed(tag, word, N, M, list_file, out_hmm, stack_unlimited=False):
    """Runs the reference trainer and returns the wall-clock duration of each of its EM iterations (E-step over the
    whole list + M-step; the initial-model builder before the first iteration is not in any of them).  The trainer
    announces every iteration on stdout ("Starting training sequence", T-FS:271) and the end of the loop ("Final
    Probability", T-FS:359); its stdout is attached to a pseudo-terminal so that libc flushes every line as it is
    written, and the lines are time-stamped as they arrive."""
    import pty
    import time
    exe = os.path.join(REF_DIR, "hmm_fs_%s" % tag)
    cmd = "%s %s %d 1 %d %s %s" % (exe, word, N, M, list_file, out_hmm)
    if stack_unlimited:
        cmd = "ulimit -s unlimited; " + cmd
    master, slave = pty.openpty()
    p = subprocess.Popen(["bash", "-c", cmd], stdout=slave, stderr=subprocess.DEVNULL, stdin=subprocess.DEVNULL, close_fds=True)
    os.close(slave)
    starts, final, tail = [], None, b""
    while True:
        try:
            chunk = os.read(master, 65536)
        except OSError:
            break
        if not chunk:
            break
        now = time.perf_counter()
        data = tail + chunk
        for _ in range(data.count(b"Starting training sequence")):
            starts.append(now)
        if b"Final Probability" in data and final is None:
            final = now
        tail = data[-32:] if b"Starting training sequence" not in data[-32:] and b"Final Probability" not in data[-32:] else b""
    os.close(master)
    if p.wait() != 0 or final is None or not starts:
        raise RuntimeError("reference trainer failed: %s" % cmd)
    edges = starts + [final]
    return [edges[k + 1] - edges[k] for k in range(len(starts))]
