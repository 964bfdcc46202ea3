#!/bin/bash
# Round-end evidence run on one B200 (under gpurun): GPU test suite, both bench arms, the ncu launch list of the
# steady-state step and `ncu --set full` captures of the dominant kernels.  Everything lands in gpurun_out/<tag>_*;
# tools/summarize_profiles.py turns it into the tracked files under profiles/.
tag=${1:-prof}
what=${2:-all}
mkdir -p gpurun_out
if [[ $what == quick ]]; then
  timeout 900 python -m pytest tests/test_cnn_gpu.py tests/test_net_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/${tag}_tests.log
fi
if [[ $what == all || $what == tests ]]; then
  timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/${tag}_tests.log
fi
if [[ $what == all || $what == bench || $what == quick ]]; then
  python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
  python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
  tail -c 600 gpurun_out/${tag}_bench_reference.json
  python - <<PY
import json
for l in open("gpurun_out/${tag}_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"])
        print({k: round(v["ms_per_step"], 3) for k, v in d["kernel_ms"].items()})
        print(d.get("parity")); print(d.get("latency_batch1")); print(d.get("config4_train"))
PY
fi
if [[ $what == all || $what == ncu || $what == quick ]]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
      python tools/profile_step.py 4 > gpurun_out/${tag}_launches.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'conv_igemm_pair_fp4_kernel|gn_apply_rows_kernel' \
      --launch-skip 70 --launch-count 12 -f -o gpurun_out/${tag}_conv_gn_full python tools/profile_step.py 3 > gpurun_out/${tag}_ncu_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'dsac_score_kernel|dsac_refine_kernel|dsac_sample_kernel' \
      --launch-skip 3 --launch-count 3 -f -o gpurun_out/${tag}_dsac_full python tools/profile_step.py 3 >> gpurun_out/${tag}_ncu_full.log 2>&1
  tail -3 gpurun_out/${tag}_ncu_full.log
  # gpurun brings back at most 64 MiB: keep the raw-page CSVs, drop reports that would not fit
  for r in conv_gn_full dsac_full; do
    ncu -i gpurun_out/${tag}_${r}.ncu-rep --page raw --csv > gpurun_out/${tag}_${r}.csv 2>/dev/null
  done
  if [[ $(du -sm gpurun_out | cut -f1) -gt 55 ]]; then rm -f gpurun_out/${tag}_conv_gn_full.ncu-rep; fi
  ls -la gpurun_out/${tag}_*
fi
