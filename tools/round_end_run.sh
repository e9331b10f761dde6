set -x
python -m pytest tests -m gpu -q 2>&1 | tail -1
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_v15.json 2> gpurun_out/bench_v15.err; cut -c1-200 gpurun_out/bench_v15.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_v15.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref_v15.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v15.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_v15.log 2>&1
python bench.py --model rodent_pair --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/bench_pair_v15.json; cut -c1-160 gpurun_out/bench_pair_v15.json
python bench.py --workload ppo --steps 3 --warmup 1 > gpurun_out/bench_ppo_v15.json; cut -c1-160 gpurun_out/bench_ppo_v15.json
python bench.py --steps 30 --warmup 5 --iterations 4 --ls-iterations 4 --no-cpu-baseline | cut -c1-130
python bench.py --steps 30 --warmup 5 --iterations 6 --ls-iterations 6 --no-cpu-baseline | cut -c1-130
python -c "
import numpy as np
from brax_rodent_run_b200.env import Rodent
t=np.zeros((4,3),np.float32)
for m in ('rodent_0','rodent_pair','rodent_new'):
    e=Rodent(t,num_envs=4096,device='cuda:0',model=m); print(m,'geometry (ctas, envs/cta, passes)',e._geometry,'smem/env',e.dims.smem_bytes)
"
