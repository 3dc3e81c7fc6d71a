import sys, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from libdwbc_b200 import *
from tests.common import *
B=65536
dev=torch.device('cuda',0)
T=lambda a,dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a),dtype=dt,device=dev).contiguous()
q,qd,rng=synth_states(B,20242)
fs=synth_fstar(rng,B)
for scale,name in ((1.0,'workload'),(0.0,'fstar=0')):
    rd=configure_product(RobotDataBatch(URDF,B,0),2)
    rd.update_kinematics(T(q),T(qd)); rd.set_contact(T(np.ones((B,2),np.uint8),torch.uint8)); rd.calc_contact_constraint()
    for l,f in enumerate(fs): rd.set_task_space(l,T(scale*f),T(posture_selector()) if l==2 else None)
    rd.solve(0xF); torch.cuda.synchronize()
    rd.set_profiling(True); rd.kernel_times(True)
    for _ in range(3): out=rd.solve(0xF)
    torch.cuda.synchronize()
    kt=rd.kernel_times(True)
    it=rd.qp_iterations().cpu().numpy()
    print(name, 'hqp ms', kt['hqp_kernel'][0]/3, 'task ms', kt['task_space_kernel'][0]/3, 'frac iters==0', (it==0).mean(), 'mean iters', it.mean(), 'p50', np.median(it), 'p90', np.percentile(it,90))
