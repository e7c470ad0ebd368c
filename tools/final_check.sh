set -x
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 100 python tools/host_cost.py 400 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -c 200 gpurun_out/r2_final_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2_final_bench.json'))
for k in ('value','ms_per_step','e2e','synchronous'): print(k, d.get(k))
print(d['c4']['seconds'], d['c4']['frames_per_s'], d['node_create']['value'], d['posegraph']['seconds'])"
