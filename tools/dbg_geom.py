import numpy as np, torch, sys
sys.path.insert(0,'/root/repo')
from coarsegrainingvae_b200 import ops, synthetic
from oracle import cgvae_oracle as orc, graph_oracle as gorc
DEV='cuda'
rng = np.random.default_rng(5)
xyz = synthetic.lattice_points(400, 2.2, rng)
pairs = gorc.make_directed(gorc.radius_graph(xyz, 9.0, True))
for it in range(3):
    g = ops.build_graph(torch.as_tensor(pairs).to(DEV), 400)
    for R, cutoff in ((8, 8.5), (10, 12.0), (4, 4.0)):
        xd = torch.as_tensor(xyz).to(DEV)
        geom = ops.edge_geometry(g, xd, xd, R, cutoff)
        eid = g.eid.cpu().numpy()
        i, j = pairs[eid, 0], pairs[eid, 1]
        r = torch.from_numpy(xyz[j] - xyz[i])
        d, unit, rbf, env = orc.edge_geometry(r, R, cutoff)
        du = (geom.unit[:, :3].cpu() - unit).abs().max(1).values
        bad = (du > 1e-6).nonzero().flatten()
        print(it, R, cutoff, "n_bad", bad.numel(), "of", du.numel(), "first bad", bad[:5].tolist(), "max", float(du.max()))
        if bad.numel():
            b = int(bad[0])
            print("  slot", b, "i,j", i[b], j[b], "col", int(g.col[b]), "cuda unit", geom.unit[b].tolist(), "oracle", unit[b].tolist(), float(d[b]))
            # which receiver does the kernel think?
            rp = g.rowptr.cpu().numpy()
            print("  rowptr around", np.searchsorted(rp, b, side='right')-1, i[b])
