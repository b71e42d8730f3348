import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle, workloads as wl
import nvdiffrast_b200.torch as dr
from nvdiffrast_b200.torch import ops
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
pos, tri, col = wl.triangle_scene(); res=(256,256)
ctx = dr.RasterizeCudaContext()
rast,_ = dr.rasterize(ctx, cu(pos), cu(tri), res)
color = np.random.default_rng(14).uniform(size=rast.shape[:3]+(1,)).astype(np.float32)
topo = dr.antialias_construct_topology_hash(cu(tri))
print("wings", topo.wings.cpu().numpy())
out, (items, cnt) = ops._antialias_fwd(cu(color), rast, cu(pos), cu(tri), topo)
torch.cuda.synchronize()
n = int(cnt[0]); it = items[:n].cpu().numpy()
act = it[it[:,3]!=0]
mine = sorted((int(r[0]&0xffff), int(r[0]>>16)&0xffff, (int(r[2])>>2)&1, int(r[2])&3, (int(r[2])>>3)&1, float(np.int32(r[3]).view(np.float32))) for r in act)
oo, oit = oracle.antialias_fwd(color, rast.cpu().numpy(), pos, tri)
orc = sorted((int(r[0]), int(r[1]), int(r[3]), int(r[4]), int(r[5]), float(np.int32(r[6]).view(np.float32))) for r in oit)
print("n items", n, "active mine", len(mine), "oracle", len(orc))
sm = {m[:3]: m for m in mine}; so = {o[:3]: o for o in orc}
only_m = [sm[k] for k in sm if k not in so]; only_o = [so[k] for k in so if k not in sm]
print("only mine", only_m[:10], len(only_m)); print("only oracle", only_o[:10], len(only_o))
diff = [(sm[k], so[k]) for k in sm if k in so and (sm[k][3:5] != so[k][3:5] or abs(sm[k][5]-so[k][5])>1e-4)]
print("diff", diff[:10], len(diff))
dy = np.random.default_rng(15).normal(size=out.shape).astype(np.float32)
gc, gp = ops._antialias_grad(cu(color), rast, cu(pos), cu(tri), cu(dy), (items, cnt))
goc, gop = oracle.antialias_grad(color, rast.cpu().numpy(), pos, tri, dy, oit)
print("mine\n", gp.cpu().numpy(), "\noracle\n", gop)
# oracle grad with MY items
myit = np.array([[m[0], m[1], 0, m[2], m[3], m[4], np.float32(m[5]).view(np.int32)] for m in mine], np.int32)
_, gop2 = oracle.antialias_grad(color, rast.cpu().numpy(), pos, tri, dy, myit)
print("oracle grad with my items\n", gop2)
