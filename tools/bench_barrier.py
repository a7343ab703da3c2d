import sys; sys.path.insert(0,'.')
import numpy as np, torch, json
import azplugins_b200 as az
N=int(sys.argv[1]) if len(sys.argv)>1 else 16000000
rng=np.random.default_rng(1)
L=(N/0.8)**(1/3)
xyz=rng.uniform(-0.5,0.5,(N,3))*L
typeid=rng.integers(0,2,N).astype(np.uint32)
state=az.State(az.Box.cube(L),["A","B"],xyz,typeid=typeid,dtype=np.float32)
peak=json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
for cls,loc in ((az.external.PlanarHarmonicBarrier,0.25*L),(az.external.SphericalHarmonicBarrier,0.4*L)):
    b=cls(location=loc); b.params["A"]=dict(k=50.0,offset=0.1); b.params["B"]=dict(k=200.0,offset=-0.1)
    b.attach(state)
    for blk in (128,256):
        b.block_size=blk
        for _ in range(3): b.compute()
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): b.compute()
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/20
        gbs=56.0*N/ms/1e6
        print("%s N=%d block %d: %.4f ms  %.1f GB/s algorithmic (56 B/particle) = %.1f%% of measured HBM peak; %.2f G particle-steps/s"%(cls.__name__,N,blk,ms,gbs,100*gbs/peak,N/ms/1e6),flush=True)
