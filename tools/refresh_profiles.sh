#!/bin/bash
# usage: tools/refresh_profiles.sh <gpurun_out subdir> <tag>   (run here, after a tools/gpu_round.sh visit)
D=gpurun_out/$1; T=${2:-r02}
C=$(cut -d' ' -f2 $D/commit.txt)
python tools/ncu_summary.py $D/prof_synth.ncu-rep > profiles/${T}_ncu_synthetic_256x256x64_pool_kernel.txt
python tools/ncu_summary.py $D/prof_synth512.ncu-rep > profiles/${T}_ncu_synthetic_512x512x64_pool_kernel.txt
python tools/ncu_summary.py $D/prof_demo.ncu-rep > profiles/${T}_ncu_demo_wg_kernel.txt
python tools/ncu_summary.py $D/prof_build.ncu-rep > profiles/${T}_ncu_request_builder_kernel.txt
python tools/ncu_evidence.py synthetic-256x256x64 $D/prof_synth.ncu-rep "profiles/${T}_ncu_synthetic_256x256x64_pool_kernel.txt: ncu --set full --clock-control none of the trace kernel at commit $C" > /dev/null
python tools/ncu_evidence.py synthetic-512x512x64 $D/prof_synth512.ncu-rep "profiles/${T}_ncu_synthetic_512x512x64_pool_kernel.txt: ncu --set full --clock-control none of the trace kernel at commit $C" > /dev/null
python tools/ncu_evidence.py demo $D/prof_demo.ncu-rep "profiles/${T}_ncu_demo_wg_kernel.txt: ncu --set full --clock-control none of the trace kernel at commit $C" > /dev/null
cp $D/bench.json profiles/${T}_bench_n1.json
cp $D/bench_ref.json profiles/${T}_bench_reference_arm.json
python tools/launch_list.py $D/launches.csv > profiles/${T}_launches_bench_synthetic.txt
ncu -i $D/prof_synth.ncu-rep --page source --csv > /tmp/_src.csv 2>/dev/null && python tools/ncu_regions.py /tmp/_src.csv 268282112 > profiles/${T}_ncu_synthetic_sass_regions.txt
cp $D/memcheck.txt profiles/${T}_compute_sanitizer_memcheck.txt
cp $D/racecheck.txt profiles/${T}_compute_sanitizer_racecheck.txt
