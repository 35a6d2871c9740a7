#!/bin/bash
# End-of-round verification on one B200 (run through gpurun): GPU tests, smoke, the bench line, and the profiles that
# DESIGN.md / profiles/ cite.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py 2>gpurun_out/bench_final.err | tail -1 > gpurun_out/bench_final.json; cat gpurun_out/bench_final.json
timeout 300 python tools/profile_step.py > /dev/null 2>&1; head -24 gpurun_out/profile_step.txt | cut -c1-75,150-230
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile-only > gpurun_out/ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt; gzip -f gpurun_out/launches.csv; head -12 gpurun_out/launches_summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:resblock_kernel -c 2 -f -o gpurun_out/resblock python bench.py --profile-only > gpurun_out/ncu_resblock.log 2>&1; ls -la gpurun_out/*.ncu-rep
