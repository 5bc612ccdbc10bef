#!/bin/bash
# Round refresh on the GPU box: bench (default, extras, reference arm), ncu launch list, ncu --set full of the hot kernels.
#   gpurun --timeout 1200 -- 'bash profiles/refresh_round.sh s50'   then   python profiles/make_round_summaries.py gpurun_out/s50
P=gpurun_out/${1:-s50}
timeout 300 python bench.py > ${P}_bench.log 2> ${P}_bench.err
timeout 500 python bench.py --extras > ${P}_bench_extras.log 2> ${P}_extras.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_ref.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${P}_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:"sample_kernel|scan_kernel|select_kernel|emit_fast_kernel" -s 12 -c 8 -o ${P}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${P}_ncu2.log 2>&1
tail -c 700 ${P}_bench.log
