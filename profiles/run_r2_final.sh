#!/bin/bash
# round 2, final tree: whole GPU suite, smoke, the default bench line and the reference arm exactly as the driver runs them
O=gpurun_out/r02f; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee $O/pytest_gpu.txt
cp gpurun_out/parity_margins.json $O/parity_margins.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $O/smoke.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>$O/bench_reference.err | tail -1 | tee $O/bench_reference.json | cut -c1-300
timeout 600 python bench.py 2>$O/bench_default.err | tail -1 | tee $O/bench_default.json | cut -c1-400
tail -2 $O/bench_default.err
