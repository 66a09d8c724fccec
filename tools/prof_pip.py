import sys, time, os
sys.path.insert(0,'.')
import torch, numpy as np
from geopolars_b200 import GeoArrowArray, GeometryType, synth
from geopolars_b200 import engine as E
dev=torch.device('cuda',0)
st=torch.cuda.Stream()
n=100_000_000
with torch.cuda.stream(st):
    ctx=E.Context(0, st.cuda_stream)
    pts=torch.empty((n,2),dtype=torch.float64,device=dev)
    E.check(ctx.lib.gpl_gen_uniform_points(ctx._h,2,0,n,1000.0,pts.data_ptr()))
    xy,ro,go=synth.star_polygons(10000,100)
    polys=ctx.upload(GeoArrowArray.polygons(xy,ro,go))
    ids=torch.empty(n,dtype=torch.int32,device=dev)
    st.synchronize()
    for rep in range(3):
        t0=time.perf_counter(); idx=E.PipIndex(polys); st.synchronize(); t1=time.perf_counter()
        print("index build ms", (t1-t0)*1e3, "bytes", idx.nbytes)
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        for k in range(3):
            e0.record(st); idx.query_device(pts.data_ptr(), n, ids.data_ptr()); e1.record(st); st.synchronize()
            print("  query ms", e0.elapsed_time(e1))
        t0=time.perf_counter(); idx.free(); t1=time.perf_counter(); print("free ms",(t1-t0)*1e3)

    # per-polygon hit histogram (the N>1 step adds it)
    idx=E.PipIndex(polys)
    idx.query_device(pts.data_ptr(), n, ids.data_ptr())
    counts=torch.zeros(10000,dtype=torch.int64,device=dev)
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    for k in range(3):
        counts.zero_()
        e0.record(st); E.check(ctx.lib.gpl_join_histogram(ctx._h, ids.data_ptr(), n, counts.data_ptr(), 10000, E.GPL_DEVICE)); e1.record(st); st.synchronize()
        print("  histogram ms", e0.elapsed_time(e1), int(counts.sum().item()))
