set -x
mkdir -p gpurun_out/r4
python bench.py --steps 20 --warmup 5 > gpurun_out/r4/bench_ml20m_n1.json 2> gpurun_out/r4/bench_ml20m_n1.err
for w in ml1m ml1m_all hm hm_nn50 stream score score_hm; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r4/bench_${w}_n1.json 2> gpurun_out/r4/bench_${w}_n1.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r4/bench_ml20m_reference.json 2> gpurun_out/r4/bench_ml20m_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4/launches_ml20m.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gram_lower_kernel|gram_mirror_kernel|gram_unpermute_kernel|gram_head_tc_kernel|gh_densify_kernel|slim_solve_warp_kernel|recommend_tc_kernel|recommend_tcfix_kernel|row_sort_bitmap_kernel|entry_pos_kernel' -c 10 -f -o gpurun_out/r4/full python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r4/ncu_full.log 2>&1
ls -la gpurun_out/r4
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r4/bench_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
    except Exception as e: print(f, "ERR", e)
PY
