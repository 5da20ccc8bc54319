# final evidence of the round: GPU tests, bench lines of the workloads touched last, launch list, full ncu capture
set -x
mkdir -p gpurun_out/r6
timeout 1400 python -m pytest tests -m gpu -x -q > gpurun_out/r6/pytest_gpu.log 2>&1; tail -3 gpurun_out/r6/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r6/bench_ml20m_n1.json 2> gpurun_out/r6/bench_ml20m_n1.err
for w in stream score score_hm ml1m hm hm_nn50; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r6/bench_${w}_n1.json 2> gpurun_out/r6/bench_${w}_n1.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r6/launches_ml20m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gram_lower_kernel|gram_mirror_kernel|gram_unpermute_kernel|gram_head_tc_kernel|gh_densify_kernel|slim_solve_warp_kernel|recommend_tc_kernel|recommend_tcfix_kernel|row_sort_bitmap_kernel|entry_pos_kernel' -c 11 -f -o gpurun_out/r6/full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r6/ncu_full.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r6/bench_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("phase_ms"))
    except Exception as e: print(f, "ERR", e)
PY
