"""Race and memory checks of the host-side thread pools (feature ingest, bulk .hmm I/O): the C sources are built with
-fsanitize=thread / address together with small drivers (tests/native/) and run on ragged inputs, tiny staging windows
and the error paths (a sink that fails, a file of another width, a missing file).  CPU only; skipped where the
compiler has no sanitizer runtime."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "speech_recognition_hmm_continuous_b200", "csrc", "host")
NATIVE = os.path.join(ROOT, "tests", "native")


def _build(tmp_path, san, driver):
    exe = str(tmp_path / ("%s_%s" % (driver, san)))
    cmd = ["gcc", "-g", "-O1", "-fsanitize=" + san, "-I", os.path.join(ROOT, "include"), os.path.join(NATIVE, driver + ".c"),
           os.path.join(NATIVE, "device_stubs.c"), os.path.join(HOST, "ingest.c"), os.path.join(HOST, "modelset.c"), "-o", exe, "-lpthread", "-lm"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if p.returncode != 0:
        pytest.skip("no %s sanitizer toolchain: %s" % (san, p.stdout.decode()[-200:]))
    return exe


@pytest.mark.parametrize("san", ["thread", "address"])
def test_reader_pools_are_clean_under_sanitizers(tmp_path, san):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0", TSAN_OPTIONS="halt_on_error=0")
    scratch = tmp_path / "data"
    scratch.mkdir()
    ok = subprocess.run([_build(tmp_path, san, "pools_ok"), str(scratch)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, timeout=300)
    out = ok.stdout.decode()
    if "unexpected memory mapping" in out or "failed to allocate" in out:   # a sanitizer runtime this kernel / container cannot host
        pytest.skip(out[-200:])
    assert ok.returncode == 0 and "Sanitizer" not in out, out[-2000:]
    lines = [l for l in out.splitlines() if l.startswith("rc=")]
    assert len(lines) == 3 and all("rc=0" in l and "sum_ok=1" in l and "frames=30000 got=30000" in l for l in lines), out
    assert "batches=1 " in lines[0] and int(lines[2].split("batches=")[1].split()[0]) > 50      # 50-frame staging windows
    assert "write rc=0 read rc=0 eq=1 word=w39" in out
    err = subprocess.run([_build(tmp_path, san, "pools_errors"), str(scratch)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, timeout=300)
    out = err.stdout.decode()
    assert err.returncode == 0 and "Sanitizer" not in out, out[-2000:]
    assert "sink failure: rc=3" in out and "wrong width: rc=5 bad=120" in out and "missing: rc=5 bad=120" in out


def test_host_model_code_is_clean_under_asan_ubsan(tmp_path):
    """Initial-model builder (M = 1..7: doubling, partial splits), host M-step, one- and two-stream .hmm files."""
    exe = str(tmp_path / "host_model")
    cmd = ["gcc", "-g", "-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-DWITH_HMM_HOST", "-I", os.path.join(ROOT, "include"),
           os.path.join(NATIVE, "host_model.c"), os.path.join(NATIVE, "device_stubs.c"), os.path.join(HOST, "hmm_host.c"),
           os.path.join(HOST, "ingest.c"), os.path.join(HOST, "modelset.c"), "-o", exe, "-lpthread", "-lm"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if p.returncode != 0:
        pytest.skip("no address/undefined sanitizer toolchain: %s" % p.stdout.decode()[-200:])
    scratch = tmp_path / "data"
    scratch.mkdir()
    r = subprocess.run([exe, str(scratch)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    out = r.stdout.decode()
    assert r.returncode == 0 and "host model code: bad=0" in out and "Sanitizer" not in out and "runtime error" not in out, out[-2000:]
