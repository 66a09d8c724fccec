import sys, time
sys.path.insert(0,'.')
import torch, numpy as np
from geopolars_b200 import GeoArrowArray, synth
from geopolars_b200 import engine as E
dev=torch.device('cuda',0)
st=torch.cuda.Stream()
n=100_000_000
def run(tag, ctx, polys, pts, ids, reps=3):
    idx=E.PipIndex(polys)
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    ts=[]
    for k in range(reps):
        e0.record(st); idx.query_device(pts.data_ptr(), pts.shape[0], ids.data_ptr()); e1.record(st); st.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{tag}: query ms {min(ts):.3f}  index MB {idx.nbytes/1e6:.1f}  hits {(ids>=0).float().mean().item():.3f}")
    idx.free()
with torch.cuda.stream(st):
    ctx=E.Context(0, st.cuda_stream)
    pts=torch.empty((n,2),dtype=torch.float64,device=dev)
    E.check(ctx.lib.gpl_gen_uniform_points(ctx._h,2,0,n,1000.0,pts.data_ptr()))
    ids=torch.empty(n,dtype=torch.int32,device=dev)
    xy,ro,go=synth.star_polygons(10000,100)
    polys=ctx.upload(GeoArrowArray.polygons(xy,ro,go))
    st.synchronize()
    run("baseline 10k polys random pts", ctx, polys, pts, ids)
    # sorted by cell
    cell=(pts[:,1]/10).floor().long()*100+(pts[:,0]/10).floor().long()
    order=torch.argsort(cell)
    spts=pts[order].contiguous(); del order, cell
    st.synchronize()
    run("10k polys, points sorted by cell", ctx, polys, spts, ids)
    del spts
    # small index: 2500 polys on 50x50 grid, points scaled to 500
    xy2,ro2,go2=synth.star_polygons(2500,50)
    polys2=ctx.upload(GeoArrowArray.polygons(xy2,ro2,go2))
    p2=(pts*0.5).contiguous()
    st.synchronize()
    run("2.5k polys (17 MB index), random pts", ctx, polys2, p2, ids)
    xy3,ro3,go3=synth.star_polygons(400,20)
    polys3=ctx.upload(GeoArrowArray.polygons(xy3,ro3,go3))
    p3=(pts*0.2).contiguous()
    st.synchronize()
    run("400 polys (2.7 MB index), random pts", ctx, polys3, p3, ids)
    # all points outside every bbox: pure streaming floor
    p4=(pts+5000.0).contiguous()
    st.synchronize()
    run("all points outside grid (stream floor)", ctx, polys, p4, ids)
