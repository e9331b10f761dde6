set -x
python -m pytest tests -q -m gpu 2>&1 | tail -2
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; cut -c1-220 gpurun_out/bench_r02_final.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r02_reference.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 30 --warmup 5 --iterations 4 --ls-iterations 4 --no-cpu-baseline --no-extra | cut -c1-130
python bench.py --steps 30 --warmup 5 --iterations 6 --ls-iterations 6 --no-cpu-baseline --no-extra | cut -c1-130
python bench.py --steps 30 --warmup 5 --model rodent_new --no-cpu-baseline --no-extra | cut -c1-130
python bench.py --steps 30 --warmup 5 --model rodent_optimized --no-cpu-baseline --no-extra | cut -c1-130
RR_WPB=1 python tools/phase_profile.py --envs 148 > gpurun_out/phase_r02_final_wpb1.txt 2>&1
python tools/phase_dump.py > /dev/null 2>&1
