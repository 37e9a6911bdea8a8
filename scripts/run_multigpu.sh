set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29601 bench.py --gpus 8 --steps 200 --warmup 20 > gpurun_out/r02_scale8.json 2> gpurun_out/r02_scale8.err; tail -2 gpurun_out/r02_scale8.err
$TR --nproc-per-node 4 --master-port 29602 bench.py --gpus 4 --workload cfg4 > gpurun_out/r02_cfg4_n4.json 2>> gpurun_out/r02_scale8.err
$TR --nproc-per-node 8 --master-port 29603 bench.py --gpus 8 --workload cfg4 > gpurun_out/r02_cfg4_n8.json 2>> gpurun_out/r02_scale8.err
$TR --nproc-per-node 8 --master-port 29604 bench.py --gpus 8 --workload cfg4 --gop 12 > gpurun_out/r02_cfg4_n8_gop12.json 2>> gpurun_out/r02_scale8.err
$TR --nproc-per-node 2 --master-port 29605 bench.py --gpus 2 --workload cfg3 --steps 300 > gpurun_out/r02_cfg3_n2.json 2>> gpurun_out/r02_scale8.err
$TR --nproc-per-node 8 --master-port 29606 bench.py --gpus 8 --workload cfg3 --steps 300 > gpurun_out/r02_cfg3_n8.json 2>> gpurun_out/r02_scale8.err
$TR --nproc-per-node 2 --master-port 29607 bench.py --gpus 2 --workload dropin --train > gpurun_out/r02_dropin_train_n2.json 2>> gpurun_out/r02_scale8.err
$TR --nproc-per-node 8 --master-port 29608 bench.py --gpus 8 --workload dropin --train > gpurun_out/r02_dropin_train_n8.json 2>> gpurun_out/r02_scale8.err
$TR --nproc-per-node 8 --master-port 29609 bench.py --gpus 8 --workload cfg5 --steps 100 --no-e2e > gpurun_out/r02_cfg5_n8.json 2>> gpurun_out/r02_scale8.err
for n in 1 2 4 8; do DSVC_PROBE_NUMA=1 $TR --nproc-per-node $n --master-port 2961$n scripts/probe/pcie_probe.py >> gpurun_out/r02_pcie_probe.jsonl 2>> gpurun_out/r02_scale8.err; done
$TR --nproc-per-node 8 --master-port 29620 scripts/probe/pcie_probe.py >> gpurun_out/r02_pcie_probe.jsonl 2>> gpurun_out/r02_scale8.err
tail -5 gpurun_out/r02_scale8.err
for f in r02_scale8 r02_cfg4_n4 r02_cfg4_n8 r02_cfg4_n8_gop12 r02_cfg3_n2 r02_cfg3_n8 r02_dropin_train_n2 r02_dropin_train_n8 r02_cfg5_n8; do echo "== $f"; head -c 600 gpurun_out/$f.json; echo; done
cat gpurun_out/r02_pcie_probe.jsonl
nproc; numactl -H 2>/dev/null | head -5
