"""Is the frame's DAG overhead a CUDA-graph effect?  The same launches issued eagerly on 4 streams
(feature warp | 3-ch warps | mv entropy | res entropy, joined by events) vs the graph replays."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from deepsvc_b200 import _lib, synthetic
from deepsvc_b200.hotpath import PFrameHotPath

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
_lib.load()
gpu_in = synthetic.to_device(synthetic.make_pframe_inputs(B=1, H=1088, W=1920, seed=16), dev)
models = bench.build_models(dev)
hp = PFrameHotPath(gpu_in, models, fuse_frame_warp=True)
streams = {b: torch.cuda.Stream(dev) for b in ("feature", "frames", "mv", "res")}


def timed(fn, n=300, label=""):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{label:50s} {e0.elapsed_time(e1) / n * 1e3:8.1f} us/frame", flush=True)


timed(lambda: hp.run_dag(streams, wide=False), label="eager, 4 streams + event joins")
timed(hp.run, label="eager, one stream (serial order)")
hp.capture(dag=True)
timed(lambda: hp.replay("branches4"), label="graph, 4 branches")
hp.capture(dag="wide")
timed(lambda: hp.replay("wide"), label="graph, one branch per launch")
hp.capture(dag=False)
timed(lambda: hp.replay("serial"), label="graph, serial order")

# when does each branch finish, relative to the fork?  (eager, 4 streams)
import statistics
main = torch.cuda.current_stream(dev)
rows = []
for it in range(30):
    fork = torch.cuda.Event(enable_timing=True)
    fork.record(main)
    ends = {}
    with torch.cuda.device(dev):
        for bname, st in streams.items():
            st.wait_event(fork)
            for fn, args, name in hp._calls:
                if hp._branch_of(name) == bname:
                    fn(*args, st.cuda_stream)
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(st)
            ends[bname] = ev
        for ev in ends.values():
            main.wait_event(ev)
    torch.cuda.synchronize()
    if it >= 5:
        rows.append({b: fork.elapsed_time(ev) * 1e3 for b, ev in ends.items()})
for b in streams:
    print(f"branch {b:8s} finishes {statistics.median(r[b] for r in rows):7.1f} us after the fork", flush=True)
