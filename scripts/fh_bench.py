import faulthandler, sys, runpy
faulthandler.dump_traceback_later(70, exit=True)
sys.argv = ["bench.py", "--steps", "5", "--warmup", "3", "--no-cpu"]
runpy.run_path("bench.py", run_name="__main__")
