import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/scripts")
sys.argv = ["check_bwd.py"]
import importlib.util
spec = importlib.util.spec_from_file_location("cb", "/root/repo/scripts/check_bwd.py")
cb = importlib.util.module_from_spec(spec)
src = open("/root/repo/scripts/check_bwd.py").read().split('if __name__ == "__main__":')[0]
exec(compile(src, "check_bwd", "exec"), cb.__dict__)
cb.timeit((1, 64, 2176, 3840), "smooth", ("auto", "staged", "direct"), n=5)
cb.timeit((1, 64, 256, 448), "smooth", ("auto", "staged", "direct"), n=10)
cb.timeit((2, 64, 544, 960), "smooth", ("auto", "staged", "direct"), n=10)
