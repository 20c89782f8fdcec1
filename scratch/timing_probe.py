import time, numpy as np, torch, sys
sys.path.insert(0,'/root/repo')
from veloslam_b200 import capi, synth
n=1<<20
pk,t=synth.hdl64_stream_tiled(n); b=synth.as_bytes(pk)
poses=synth.ins_trajectory(31000)
ctx=capi.Context(0,max_batch_packets=n,max_poses=40000,n_slots=1)
ctx.set_calibration(synth.calib_hdl64()); ctx.set_poses(*poses)
d_pk=torch.from_numpy(b).cuda(); d_t=torch.from_numpy(t).cuda(); torch.cuda.synchronize()
for i in range(3):
    r=ctx.wait(ctx.submit(d_pk,d_t,n=n,stride=1206,flags=1,t_base_us=int(t[0])),frames=False)
ts=[];tw=[]
for i in range(10):
    a=time.perf_counter(); tk=ctx.submit(d_pk,d_t,n=n,stride=1206,flags=1,t_base_us=int(t[0])); b1=time.perf_counter()
    r=ctx.wait(tk,frames=False); c=time.perf_counter()
    ts.append(b1-a); tw.append(c-b1)
print('submit ms',np.mean(ts)*1e3,'wait ms',np.mean(tw)*1e3,'gpu_ms',r.gpu_ms,'decode',r.decode_ms)
import ctypes as C
L=capi.load_library()
res=capi.Result()
tw2=[]
for i in range(10):
    tk=ctx.submit(d_pk,d_t,n=n,stride=1206,flags=1,t_base_us=int(t[0]))
    a=time.perf_counter(); L.vs_wait(ctx._h,tk,C.byref(res)); c=time.perf_counter(); tw2.append(c-a)
print('raw wait ms',np.mean(tw2)*1e3)
