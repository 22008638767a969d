import json, sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from conftest import load_problem
from adpres_b200 import capi
ref = json.load(open('/root/repo/tests/golden/' + (sys.argv[1] if len(sys.argv) > 1 else 'c2_oracle_result.json')))
p = load_problem("IAEA3Ds").refine(xdiv=[10] + [20] * 8, ydiv=[20] * 8 + [10], zdiv=[ref['zdiv']] * 19)
s = capi.Solver(p, nin=10, nac=5, nupd=50, nout=3000); s.enable_trace(); rc, n = s.outer(1)
print(rc, n, s.state()['Ke'], ref['keff'])
for (q, k, ser, fer) in s.trace_rows:
    if q in (1,2,3,5,8,10,15,20,30,50,51,60,100,150,200,300,370) and q <= len(ref['trace_ke']): print(q, '%.3e'%(k-ref['trace_ke'][q-1]), '%.3e %.3e'%(ser, ref['trace_ser'][q-1]))
print(s.trace_nodal[:4]); print(ref['nodal_updates'][:4])
rc, pw = s.powdis(); asm, asm_ref = p.asm_power(pw), np.array(ref['asm_power']); nz = asm_ref > 0
print('asm power max rel diff', np.abs(asm[nz]/asm_ref[nz]-1).max())
idx = np.array(sorted(int(i) for i in ref['power_samples'])); rp = np.array([ref['power_samples'][str(i)] for i in idx]); z = rp > 1e-12
print('node power max rel diff', np.abs(pw[idx][z]/rp[z]-1).max())
