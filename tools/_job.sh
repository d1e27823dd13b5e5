python -m pytest tests/test_gpu_train.py tests/test_gpu_propagate.py -x -q 2>&1 | tail -2
python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s17_n1.json 2> gpurun_out/bench_s17_n1.err; echo "n1 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s17_n1.json').read().strip().splitlines()[-1])
print('value %.3gM ms/step %.4f e2e %.3gM (%.4f ms) spmm %.1f us eval %.3gM users/s (%.2f ms) e2e eval %.3gM'%(d['value']/1e6,d['ms_per_step'],d['e2e']['value']/1e6,d['e2e']['ms_per_step'],d['roofline']['avg_launch_ms']*1e3,d['eval']['users_per_s']/1e6,d['eval']['ms'],d['eval']['e2e_users_per_s']/1e6))
PY
python - <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np, torch, scipy.sparse as sp
from bench import make_data
from arlib_b200.graph import DeviceGraph
D=make_data("gowalla",0.5); U,I,E=D["U"],D["I"],D["E"]; N=U+I
half=sp.csr_matrix((np.ones(E,dtype=np.float32),(D["tu"],D["ti"]+U)),shape=(N,N),dtype=np.float32)
g=DeviceGraph.from_dataloader_adj(half+half.T,"cuda:0")
for d in (64,32,16,8):
    print("d",d,"autotuned segment", g.autotune(d).segment)
PY
