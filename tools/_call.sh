timeout 600 python -m pytest tests/test_policy.py -m gpu -q -s 2>&1 | grep -E "max \|diff|passed|failed|rror|fused" | cut -c1-200
for n in 1 37 300 1025; do GRP_CONV23=fused timeout 100 python - <<PY
import torch, numpy as np, os, sys
sys.path.insert(0, ".")
from mujoco_rl_manipulate_unknown_objects_b200.policy import GripperPolicy
n=$n
pol=GripperPolicy(max_envs=n, seed=3)
obs=torch.randint(0,256,(n,5,64,64),dtype=torch.uint8,device="cuda")
a=pol(obs).clone(); os.environ["GRP_CONV23"]="split"; b=pol(obs).clone()
print("n=%d fused==split:"%n, bool(torch.equal(a,b)))
PY
done
