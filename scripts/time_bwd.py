"""Timing of the 64-ch warp backward (both gradients) per algorithm; env knobs are read by the library."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.check_bwd import timeit  # noqa: E402
algos = [a for a in sys.argv[1:] if a in ("direct", "staged", "gather")] or ["staged", "gather"]
timeit((1, 64, 1088, 1920), "smooth", algos)
timeit((8, 64, 256, 256), "smooth", algos)
