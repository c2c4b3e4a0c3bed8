import os, sys, torch, time
sys.path.insert(0, '/root/repo')
from sparsebase_b200 import lib, synth
lib.load()
dev=torch.device('cuda',0)
n, rp, col, vals = synth.poisson2d(4096,4096,device=dev)
for _ in range(2): lib.rcm_reorder(n, rp, col)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(3): lib.rcm_reorder(n, rp, col)
torch.cuda.synchronize(); ms=(time.perf_counter()-t)/3*1e3
st=lib.rcm_last_stats()
print(os.environ.get('SB200_RCM_CLUSTER'), round(ms,1), {k:round(v/st['levels_narrow']) for k,v in st['phase_cycles'].items()})
