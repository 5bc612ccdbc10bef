#!/bin/bash
# Round refresh on the GPU box: bench (default incl. extras, reference arm), ncu launch list, ncu --set full of the hot kernels.
#   gpurun --timeout 1500 -- 'bash profiles/refresh_round.sh r02'   then   python profiles/make_round_summaries.py gpurun_out/r02 r02
P=gpurun_out/${1:-r02}
timeout 400 python bench.py > ${P}_bench_default.json 2> ${P}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_reference_arm.json 2> ${P}_ref.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > ${P}_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > ${P}_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:"sample_kernel|scan_native|select_kernel|emit_fast_kernel" -s 12 -c 8 -o ${P}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > ${P}_ncu2.log 2>&1
tail -c 600 ${P}_bench_default.json
