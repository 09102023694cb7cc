#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --solves > gpurun_out/solves.json 2> gpurun_out/solves.err; echo "solves rc=$?"
cat gpurun_out/solves.md; tail -3 gpurun_out/solves.err
timeout 600 python tools/graph_probe.py 32 64 128 > gpurun_out/graph_probe.md 2> gpurun_out/graph_probe.err; echo "probe rc=$?"
grep "^|" gpurun_out/graph_probe.md; nproc; lscpu | grep "Model name"
