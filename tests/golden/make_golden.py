"""Generates the committed golden fixtures from the reference checkout (/root/reference).

Run in the build container only (the GPU box has no reference checkout):
    python tests/golden/make_golden.py

Writes into tests/golden/:
  citeseer_half.npz   Citeseer inputs of the reference (data/matrices/citeseer_*.txt) rounded to
                      binary16 exactly as main_float.cpp / mmult-master.ipynb load them, plus the
                      reference's recorded outputs:
                        csim_rows / csim_vals   hls/.../csim/report/mmult_top_csim.log:21-62
                        nb37_row0               jupyter/test/mmult-master.ipynb cell 37 output
                        nb55_row0               jupyter/test/mmult-master.ipynb cell 55 output
  toy.npz             the 4x4 hand-checkable fixtures test_*.txt (main_float.cpp:102-111)
  qlayer_*.npz        inputs + outputs of the reference's OWN emulation branch
                      (demo/sgrace_lib/sgrace.py:563-681) executed here: sgrace.py is imported
                      unmodified with tiny stand-ins for its unavailable imports
                      (torch_geometric, torch_scatter) and the reference's emulation config.py.
  sym_norm2.npz       inputs + outputs of the reference's sym_norm2 (sgrace.py:18-51)
  backward_emulation.npz  saved tensors, grad_output and the three gradients of the reference's software backward
                      (FPYNQ_GAT.backward, accb = 0), GCN and GAT, from the imported sgrace.py
  ref_hls_fix16.npz   the same for the EIGHTBIT configuration (ap_fixed<16,2>): int16 codes, incl. a set that wraps around
  ref_hls_*.npz       outputs of the reference HLS source compiled natively (oracle/_ref) on
                      small seeded inputs, so the GPU box can check against the real kernel code
  real_mol.npz        the reference's own molecule matrices (data/matrices/mol_*.txt, main_float.cpp:40-51:
                      N 2273, M 7, nnz 5028 / 2273) and the two-layer forward of BASELINE configs[0]
                      (layer 1 sparse + ReLU, layer 2 dense gemm_mode on layer 1's output) computed by the
                      reference HLS source (oracle/_ref), HALF and FLOAT builds, hidden 16 and 32
  real_cora.npz       data/matrices/cora_{adj,feat,weights}.txt (main_float.cpp:73-82) and the layer output of
                      the reference HLS source, hidden 16, HALF and FLOAT builds
  demo_model_cora.npz the reference's demo model (GAT_PYNQ: att2 -> Relu_SGRACE -> conv22 -> Linear) composed from the
                      reference's unmodified sgrace.py pieces, emulation mode, on the real Cora files: seeded
                      parameters, both layer outputs and the logits, 8- and 4-bit, GCN and GAT; and the state_dict keys
                      of demo/zcu104/model_Photo_8bit.ptx
  real_mutag.npz      jupyter/molecule_gcn/MUTAG/raw/* (the notebook's dataset: 188 graphs, 3371 nodes, 7442
                      edges, 7 one-hot node labels) and the notebook's two GCN layers (7 -> 64 -> 64, unnormalised
                      0/1 adjacency without self-loops, cells 16-18) computed by the reference HLS source
"""
import os
import re
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
MAT = REF + "/gnn-rfsoc-mt-all-2022/data/matrices/"

from oracle import oracle as O  # noqa: E402


def citeseer():
    adj = O.load_csr_txt(MAT + "citeseer_adj.txt")
    fea = O.load_csr_txt(MAT + "citeseer_feat.txt")
    w = O.load_dense_txt(MAT + "citeseer_weights.txt")
    log = open(REF + "/gnn-rfsoc-mt-all-2022/hls/gnn/solution1/gnn/solution1/csim/report/mmult_top_csim.log").read()
    rows, cols, vals = [], [], []
    for m in re.finditer(r"out :data index= (\d+) (\d+) kernel = (\S+)", log):
        rows.append(int(m.group(1)))
        cols.append(int(m.group(2)))
        vals.append(float(m.group(3)))
    assert len(vals) == 42
    import json
    nb = json.load(open(REF + "/jupyter/test/mmult-master.ipynb"))
    out37 = "".join(nb["cells"][37]["outputs"][0]["text"])
    out55 = "".join(nb["cells"][55]["outputs"][0]["text"])
    nb37 = np.array(re.findall(r"-?\d+\.\d+(?:e-?\d+)?", out37), dtype=np.float64)
    nb55 = np.array(re.findall(r"-?\d+\.\d+(?:e-?\d+)?", out55), dtype=np.float64)
    assert len(nb37) == 16 and len(nb55) == 21
    assert np.all(fea.val == 1.0)
    np.savez_compressed(
        os.path.join(HERE, "citeseer_half.npz"),
        adj_rowptr=adj.rowptr, adj_col=adj.col.astype(np.uint16), adj_val_f16=adj.val.astype(np.float16),
        fea_rowptr=fea.rowptr, fea_col=fea.col.astype(np.uint16),   # all feature values are 1.0
        w_f16=w.astype(np.float16),
        csim_rows=np.array(rows, np.int32), csim_cols=np.array(cols, np.int32),
        csim_vals=np.array(vals, np.float64), nb37_row0=nb37, nb55_row0=nb55)
    print("citeseer_half.npz", adj.n, adj.nnz, fea.nnz, w.shape)


def toy():
    d = {}
    for name in ("test_adj", "test_adj2", "test_feat", "test_feat2"):
        c = O.load_csr_txt(MAT + name + ".txt")
        d[name + "_rowptr"], d[name + "_col"], d[name + "_val"] = c.rowptr, c.col, c.val
    for name in ("test_weights", "test_weights2"):
        d[name] = O.load_dense_txt(MAT + name + ".txt")
    np.savez_compressed(os.path.join(HERE, "toy.npz"), **d)
    print("toy.npz")


# ---------------------------------------------------------------------------------------
# import the reference's sgrace.py unmodified
# ---------------------------------------------------------------------------------------
def import_reference_sgrace(w_qbits, compute_attention):
    def degree(index, num_nodes=None, dtype=None):
        n = int(index.max()) + 1 if num_nodes is None else num_nodes
        out = torch.zeros(n, dtype=dtype or torch.get_default_dtype())
        return out.scatter_add_(0, index, torch.ones(index.numel(), dtype=out.dtype))

    def add_remaining_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
        n = num_nodes
        mask = edge_index[0] != edge_index[1]
        loop_index = torch.arange(0, n, dtype=edge_index.dtype).unsqueeze(0).repeat(2, 1)
        loop_attr = None
        if edge_attr is not None:
            loop_attr = edge_attr.new_full((n,), fill_value)
            inv = ~mask
            loop_attr[edge_index[0][inv]] = edge_attr[inv]      # keep existing self-loop weights
            edge_attr = torch.cat([edge_attr[mask], loop_attr], dim=0)
        edge_index = torch.cat([edge_index[:, mask], loop_index], dim=1)
        return edge_index, edge_attr

    def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
        n = num_nodes
        loop_index = torch.arange(0, n, dtype=edge_index.dtype).unsqueeze(0).repeat(2, 1)
        if edge_attr is not None:
            edge_attr = torch.cat([edge_attr, edge_attr.new_full((n,), fill_value)], dim=0)
        return torch.cat([edge_index, loop_index], dim=1), edge_attr

    def sort_edge_index(edge_index, edge_attr=None, num_nodes=None):
        n = int(edge_index.max()) + 1
        key = edge_index[0] * n + edge_index[1]
        perm = torch.argsort(key, stable=True)
        return edge_index[:, perm], (None if edge_attr is None else edge_attr[perm])

    def scatter_add(src, index, dim=0, dim_size=None):
        out = torch.zeros(dim_size, dtype=src.dtype)
        return out.scatter_add_(0, index, src)

    tg = types.ModuleType("torch_geometric")
    tg.transforms = types.ModuleType("torch_geometric.transforms")
    tg.utils = types.ModuleType("torch_geometric.utils")
    tg.utils.add_remaining_self_loops = add_remaining_self_loops
    tg.utils.add_self_loops = add_self_loops
    tg.utils.sort_edge_index = sort_edge_index
    tg.utils.degree = degree
    ts = types.ModuleType("torch_scatter")
    ts.scatter_add = scatter_add
    sys.modules.update({"torch_geometric": tg, "torch_geometric.transforms": tg.transforms,
                        "torch_geometric.utils": tg.utils, "torch_scatter": ts})
    for m in ("config", "sgrace"):
        sys.modules.pop(m, None)
    sys.path.insert(0, REF + "/demo/emulation")
    sys.path.insert(0, REF + "/demo/sgrace_lib")
    import config
    config.acc = 0
    config.accb = 0
    config.fake_quantization = 1
    config.w_qbits = w_qbits
    config.compute_attention = compute_attention
    config.profiling = 0
    config.show_max_min = 0
    import sgrace
    sgrace.init_SGRACE()
    sys.path.pop(0)
    sys.path.pop(0)
    return config, sgrace


def small_graph(rng, n, m, p, dens_x=0.2, extra_edges=3):
    rows = np.concatenate([np.arange(n)] * extra_edges)
    cols = rng.integers(0, n, size=len(rows))
    r = np.concatenate([rows, cols])
    c = np.concatenate([cols, rows])
    keep = r != c
    ei = np.unique(np.stack([r[keep], c[keep]]), axis=1)
    x = (rng.random((n, m)) < dens_x) * rng.random((n, m))
    x[rng.integers(0, n, 3)] = 0.0          # a few all-zero feature rows
    w = rng.uniform(-1.2, 1.2, size=(m, p)) * (rng.random((m, p)) < 0.9)   # some clipping, some zeros
    att = rng.uniform(-1.0, 1.0, size=(2 * p, 1))
    return ei.astype(np.int64), x.astype(np.float32), w.astype(np.float32), att.astype(np.float32)


def qlayers():
    rng = np.random.default_rng(7)
    ei, x, w, att = small_graph(rng, 96, 40, 16)
    n = x.shape[0]
    for qbits in (8, 4, 2, 1):
        for gat in (0, 1):
            config, sg = import_reference_sgrace(qbits, gat)
            # the model's own preprocessing: sym_norm2 (sgrace.py:18-51, demo_sgrace.py:233-267)
            edge_index, norm = sg.sym_norm2(torch.from_numpy(ei), n, None, 1, torch.float32)
            adj = torch.sparse_coo_tensor(edge_index, norm, (n, n))
            layer = sg.GATConv_SGRACE(x.shape[1], w.shape[1], nheads=1, bias=False)
            layer.weight.data = torch.from_numpy(w.copy())
            layer.attention.data = torch.from_numpy(att.copy())
            outs = {}
            for relu in (0, 1):
                for dense in (0, 1):
                    with torch.no_grad():
                        out = layer.forward(gat, dense, relu, torch.from_numpy(x.copy()), edge_index, norm, adj)
                    outs[f"out_relu{relu}_dense{dense}"] = out.numpy().astype(np.float32)
            consts = {k: float(getattr(sg, k)) for k in ("w_s", "a_s", "f_s", "deq_o")}
            consts.update({k: int(getattr(sg, k)) for k in ("w_z", "a_z", "f_z", "scale_fea", "internal_quantization")})
            np.savez_compressed(
                os.path.join(HERE, f"qlayer_q{qbits}_gat{gat}.npz"),
                edge_index=edge_index.numpy().astype(np.int32), norm=norm.numpy().astype(np.float32),
                x=x, w=w, attention=att, qbits=qbits, gat=gat,
                **{"c_" + k: np.array(v) for k, v in consts.items()}, **outs)
            print(f"qlayer_q{qbits}_gat{gat}.npz", consts, {k: float(np.abs(v).max()) for k, v in outs.items()})
    # sym_norm2 fixture (pre-processing, "next" row)
    config, sg = import_reference_sgrace(8, 0)
    e2, n2 = sg.sym_norm2(torch.from_numpy(ei), n, None, 1, torch.float32)
    np.savez_compressed(os.path.join(HERE, "sym_norm2.npz"), edge_index_in=ei.astype(np.int32), num_nodes=n,
                        fill=1, edge_index_out=e2.numpy().astype(np.int32), norm_out=n2.numpy().astype(np.float32))
    print("sym_norm2.npz", e2.shape)



def backward_emulation():
    """Software backward of FPYNQ_GAT (sgrace.py:880-1110, accb = 0) run by the imported reference: the tensors its
    forward saved, a seeded grad_output and the three gradients it returns, GCN and GAT."""
    rng = np.random.default_rng(7)
    ei, x, w, att = small_graph(rng, 96, 40, 16)
    n = x.shape[0]
    d = {}
    for gat in (0, 1):
        config, sg = import_reference_sgrace(8, gat)
        edge_index, norm = sg.sym_norm2(torch.from_numpy(ei), n, None, 1, torch.float32)
        adj = torch.sparse_coo_tensor(edge_index, norm, (n, n))
        layer = sg.GATConv_SGRACE(x.shape[1], w.shape[1], nheads=1, bias=False)
        layer.weight.data = torch.from_numpy(w.copy())
        layer.attention.data = torch.from_numpy(att.copy())
        xt = torch.from_numpy(x.copy()).requires_grad_(True)
        out = layer.forward(gat, 0, 1, xt, edge_index, norm, adj)
        node = out.grad_fn
        assert "FPYNQ_GAT" in type(node).__name__
        saved = node.saved_tensors                      # adj, input, weights, e, attentions, output
        names = ("adj", "input", "weights", "e", "attentions", "output")
        for nm, t in zip(names, saved):
            if nm == "adj":
                continue
            d[f"gat{gat}_{nm}"] = t.detach().to_dense().numpy().astype(np.float32) if t.is_sparse else t.detach().numpy().astype(np.float32)
        g = torch.from_numpy(np.random.default_rng(11 + gat).standard_normal(tuple(out.shape)).astype(np.float32))
        out.backward(g)
        d[f"gat{gat}_grad_output"] = g.numpy()
        d[f"gat{gat}_grad_input"] = xt.grad.numpy().astype(np.float32)
        d[f"gat{gat}_grad_weights"] = layer.weight.grad.numpy().astype(np.float32)
        d[f"gat{gat}_grad_attention"] = layer.attention.grad.numpy().astype(np.float32)
        d[f"gat{gat}_alpha"] = np.float32(node.alpha) if hasattr(node, "alpha") else np.float32(0.2)
        d["edge_index"] = edge_index.numpy().astype(np.int32)
        d["norm"] = norm.numpy().astype(np.float32)
    d["n"] = n
    np.savez_compressed(os.path.join(HERE, "backward_emulation.npz"), **d)
    print("backward_emulation.npz", {k: v.shape for k, v in d.items() if hasattr(v, "shape") and "grad" in k})

def ref_hls():
    """Outputs of the reference HLS kernel source (compiled in oracle/_ref) on seeded inputs."""
    if not (O.ref_available("half") and O.ref_available("float")):
        print("oracle/_ref not built; run `make -C oracle ref` first")
        return
    from sgracex1_b200 import graphs as G
    rng = np.random.default_rng(11)
    p = G.cora_shape(seed=3, P=16, n=600, m=300, nnz_adj=3000, nnz_fea=6000)
    p.fea_val = rng.uniform(0.0, 1.0, size=len(p.fea_val)).astype(np.float32)
    xd = (rng.random((p.N, 24)) < 0.5) * rng.uniform(-1, 1, size=(p.N, 24))
    wd = rng.uniform(-0.5, 0.5, size=(24, 10)).astype(np.float32)
    d = dict(N=p.N, M=p.M, adj_rowptr=p.adj_rowptr, adj_col=p.adj_col, adj_val=p.adj_val,
             fea_rowptr=p.fea_rowptr, fea_col=p.fea_col, fea_val=p.fea_val, W=p.W, x_dense=xd.astype(np.float32),
             w_dense=wd)
    for kind, dt in (("half", O.F16), ("float", O.F32)):
        adj = (p.adj_rowptr, p.adj_col, O.to_storage(p.adj_val, dt))
        fea = (p.fea_rowptr, p.fea_col, O.to_storage(p.fea_val, dt))
        for P in (16, 7):
            for relu in (0, 1):
                B = O.to_storage(O.weights_to_B(p.W[:, :P]), dt)
                d[f"{kind}_sparse_P{P}_relu{relu}"] = O.ref_layer(kind=kind, N=p.N, M_fea=p.M, P=P, adj=adj, B=B,
                                                                  fea=fea, relu=relu)
        Bd = O.to_storage(O.weights_to_B(wd), dt)
        d[f"{kind}_dense_P10_relu1"] = O.ref_layer(kind=kind, N=p.N, M_fea=24, P=10, adj=adj, B=Bd,
                                                   x_dense=O.to_storage(xd, dt), relu=1)
    np.savez_compressed(os.path.join(HERE, "ref_hls.npz"), **d)
    print("ref_hls.npz", sorted(k for k in d if "relu" in k))
    ref_hls_fix16()


def ref_hls_fix16():
    """The same source in its EIGHTBIT configuration (ap_fixed<16,2>, latency 1; oracle/_ref/libsgrace_hlsref_fix16.so):
    int16 Q2.14 outputs on seeded inputs, one set small enough to stay in range and one that wraps around."""
    if not O.ref_available("fix16"):
        print("oracle/_ref fix16 build missing; run `make -C oracle ref` first")
        return
    from sgracex1_b200 import graphs as G
    d = {}
    for tag, scale in (("small", 0.5), ("wrap", 3.0)):
        rng = np.random.default_rng(21 if tag == "small" else 22)
        p = G.cora_shape(seed=5, P=16, n=400, m=120, nnz_adj=2000, nnz_fea=3600)
        av = (p.adj_val * scale * 2).astype(np.float32)
        fv = rng.uniform(-scale, scale, size=len(p.fea_val)).astype(np.float32)
        W = rng.uniform(-scale, scale, size=(p.M, 16)).astype(np.float32)
        xd = ((rng.random((p.N, 24)) < 0.5) * rng.uniform(-scale, scale, size=(p.N, 24))).astype(np.float32)
        wd = rng.uniform(-0.5, 0.5, size=(24, 10)).astype(np.float32)
        adj = (p.adj_rowptr, p.adj_col, O.to_storage(av, O.FIX16))
        fea = (p.fea_rowptr, p.fea_col, O.to_storage(fv, O.FIX16))
        d.update({f"{tag}_N": p.N, f"{tag}_M": p.M, f"{tag}_adj_rowptr": p.adj_rowptr, f"{tag}_adj_col": p.adj_col,
                  f"{tag}_adj_val": adj[2], f"{tag}_fea_rowptr": p.fea_rowptr, f"{tag}_fea_col": p.fea_col, f"{tag}_fea_val": fea[2],
                  f"{tag}_x_dense": O.to_storage(xd, O.FIX16)})
        for P in (16, 7):
            B = O.to_storage(O.weights_to_B(W[:, :P]), O.FIX16)
            d[f"{tag}_B_P{P}"] = B
            for relu in (0, 1):
                d[f"{tag}_sparse_P{P}_relu{relu}"] = O.ref_layer(kind="fix16", N=p.N, M_fea=p.M, P=P, adj=adj, B=B, fea=fea, relu=relu)
        Bd = O.to_storage(O.weights_to_B(wd), O.FIX16)
        d[f"{tag}_B_dense"] = Bd
        d[f"{tag}_dense_P10_relu1"] = O.ref_layer(kind="fix16", N=p.N, M_fea=24, P=10, adj=adj, B=Bd, x_dense=d[f"{tag}_x_dense"], relu=1)
    np.savez_compressed(os.path.join(HERE, "ref_hls_fix16.npz"), **d)
    print("ref_hls_fix16.npz", sorted(k for k in d if "relu" in k))


def _kinds():
    return (("half", O.F16), ("float", O.F32))


def real_mol():
    adj = O.load_csr_txt(MAT + "mol_adj.txt")
    fea = O.load_csr_txt(MAT + "mol_feat.txt")
    xd = O.load_dense_txt(MAT + "mol_feat_dense.txt")
    w = O.load_dense_txt(MAT + "mol_weights.txt")
    N, M = adj.n, xd.shape[1]
    assert (N, adj.nnz, fea.nnz, M) == (2273, 5028, 2273, 7) and w.shape[0] == 7    # main_float.cpp:41-45
    # the dense file is the same matrix as the CSR one
    chk = np.zeros((N, M), np.float32)
    for r in range(N):
        chk[r, fea.col[fea.rowptr[r]:fea.rowptr[r + 1]]] = fea.val[fea.rowptr[r]:fea.rowptr[r + 1]]
    assert np.array_equal(chk, xd)
    d = dict(adj_rowptr=adj.rowptr, adj_col=adj.col.astype(np.uint16), adj_val=adj.val,
             fea_rowptr=fea.rowptr, fea_col=fea.col.astype(np.uint8), fea_val=fea.val, w=w)
    rng = np.random.default_rng(12345)
    for P in (16, 32):
        w2 = rng.uniform(-1.0 / np.sqrt(P), 1.0 / np.sqrt(P), size=(P, P)).astype(np.float32)   # notebook cell 17:53-55
        d[f"w2_P{P}"] = w2
        for kind, dt in _kinds():
            a = (adj.rowptr, adj.col, O.to_storage(adj.val, dt))
            f = (fea.rowptr, fea.col, O.to_storage(fea.val, dt))
            B1 = O.to_storage(O.weights_to_B(w[:, :P]), dt)
            h1 = O.ref_layer(kind=kind, N=N, M_fea=M, P=P, adj=a, B=B1, fea=f, relu=1)
            h1d = O.ref_layer(kind=kind, N=N, M_fea=M, P=P, adj=a, B=B1, x_dense=O.to_storage(xd, dt), relu=1)
            B2 = O.to_storage(O.weights_to_B(w2), dt)
            h2 = O.ref_layer(kind=kind, N=N, M_fea=P, P=P, adj=a, B=B2, x_dense=h1, relu=0)
            d[f"{kind}_P{P}_layer1"], d[f"{kind}_P{P}_layer1_dense"], d[f"{kind}_P{P}_layer2"] = h1, h1d, h2
    np.savez_compressed(os.path.join(HERE, "real_mol.npz"), **d)
    print("real_mol.npz", N, adj.nnz, sorted(k for k in d if "layer" in k))


def real_cora():
    adj = O.load_csr_txt(MAT + "cora_adj.txt")
    fea = O.load_csr_txt(MAT + "cora_feat.txt")
    w = O.load_dense_txt(MAT + "cora_weights.txt")
    N, M, P = adj.n, w.shape[0], 16
    assert (N, M, adj.nnz, fea.nnz) == (2708, 1433, 13264, 49216)                  # main_float.cpp:74-78
    fea_one = bool(np.all(fea.val == 1.0))
    d = dict(adj_rowptr=adj.rowptr, adj_col=adj.col.astype(np.uint16), adj_val=adj.val,
             fea_rowptr=fea.rowptr, fea_col=fea.col.astype(np.uint16), w=w[:, :P].copy(), fea_all_ones=fea_one)
    if not fea_one:
        d["fea_val"] = fea.val
    for kind, dt in _kinds():
        a = (adj.rowptr, adj.col, O.to_storage(adj.val, dt))
        f = (fea.rowptr, fea.col, O.to_storage(fea.val, dt))
        B = O.to_storage(O.weights_to_B(w[:, :P]), dt)
        for relu in (0, 1):
            d[f"{kind}_relu{relu}"] = O.ref_layer(kind=kind, N=N, M_fea=M, P=P, adj=a, B=B, fea=f, relu=relu)
    np.savez_compressed(os.path.join(HERE, "real_cora.npz"), **d)
    print("real_cora.npz", N, adj.nnz, fea.nnz, "features all ones:", fea_one)


def real_mutag():
    raw = REF + "/jupyter/molecule_gcn/MUTAG/raw/"
    edges = np.loadtxt(raw + "MUTAG_A.txt", delimiter=",", dtype=np.int64) - 1        # 1-based (row, col) pairs
    indicator = np.loadtxt(raw + "MUTAG_graph_indicator.txt", dtype=np.int64) - 1
    node_labels = np.loadtxt(raw + "MUTAG_node_labels.txt", dtype=np.int64)
    graph_labels = np.loadtxt(raw + "MUTAG_graph_labels.txt", dtype=np.int64)
    N, H = len(indicator), 64
    assert (N, len(edges), len(graph_labels), int(node_labels.max()) + 1) == (3371, 7442, 188, 7)
    # the notebook's batch of all 188 graphs: dense 0/1 adjacency -> CSR (cell 18:53-56), one-hot x -> CSR (18:94-117)
    order = np.lexsort((edges[:, 1], edges[:, 0]))
    e = edges[order]
    rowptr = np.zeros(N + 1, np.int32)
    np.add.at(rowptr, e[:, 0] + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    col = e[:, 1].astype(np.int32)
    val = np.ones(len(col), np.float32)
    frp = np.arange(N + 1, dtype=np.int32)
    fci = node_labels.astype(np.int32)
    fva = np.ones(N, np.float32)
    rng = np.random.default_rng(12345)
    w1 = rng.uniform(-1.0 / np.sqrt(H), 1.0 / np.sqrt(H), size=(7, H)).astype(np.float32)
    w2 = rng.uniform(-1.0 / np.sqrt(H), 1.0 / np.sqrt(H), size=(H, H)).astype(np.float32)
    d = dict(edges=edges.astype(np.int16), graph_indicator=indicator.astype(np.int16), node_labels=node_labels.astype(np.int8),
             graph_labels=graph_labels.astype(np.int8), w1=w1, w2=w2)
    for kind, dt in _kinds():
        a = (rowptr, col, O.to_storage(val, dt))
        f = (frp, fci, O.to_storage(fva, dt))
        h1 = O.ref_layer(kind=kind, N=N, M_fea=7, P=H, adj=a, B=O.to_storage(O.weights_to_B(w1), dt), fea=f, relu=1)
        h2 = O.ref_layer(kind=kind, N=N, M_fea=H, P=H, adj=a, B=O.to_storage(O.weights_to_B(w2), dt), x_dense=h1, relu=0)
        if kind == "half":
            d["half_layer1"] = h1
        d[f"{kind}_layer2"] = h2
    np.savez_compressed(os.path.join(HERE, "real_mutag.npz"), **d)
    print("real_mutag.npz", N, len(col))


def demo_model():
    """The reference's demo model (demo/emulation/demo_sgrace.py:271-401, GAT_PYNQ) on the reference's own Cora
    files, in the emulation mode (acc = 0): the class cannot be imported (the script loads Planetoid at import), so
    its forward is composed here from the reference's own, unmodified pieces -- sym_norm2, GATConv_SGRACE,
    Relu_SGRACE from demo/sgrace_lib/sgrace.py -- in the order of demo_sgrace.py:292-399."""
    adj = O.load_csr_txt(MAT + "cora_adj.txt")
    fea = O.load_csr_txt(MAT + "cora_feat.txt")
    n, m, hidden, classes = adj.n, 1433, 16, 7
    rows = np.repeat(np.arange(n), np.diff(adj.rowptr))
    keep = rows != adj.col                               # the file holds A + I normalised; the model wants the raw edges
    ei = np.stack([rows[keep], adj.col[keep]]).astype(np.int64)
    assert ei.shape[1] == 10556
    x = np.zeros((n, m), np.float32)
    x[np.repeat(np.arange(n), np.diff(fea.rowptr)), fea.col] = fea.val
    avg_deg = ei.shape[1] / n
    fill = int(np.trunc(np.log2(avg_deg)))
    d = dict(edge_index=ei.astype(np.int32), fea_rowptr=fea.rowptr, fea_col=fea.col.astype(np.uint16), average_node_degree=avg_deg,
             state_dict_keys=np.array(["att2.weight", "att2.attention", "att2.bias", "conv22.weight", "conv22.attention",
                                       "conv22.bias", "lin.weight", "lin.bias"]))
    import torch.nn.functional as F
    for qbits in (8, 4):
        for gat in (0, 1):
            config, sg = import_reference_sgrace(qbits, gat)
            torch.manual_seed(12345)
            att2 = sg.GATConv_SGRACE(m, hidden, 1, dropout=0.1, alpha=0.2, concat=False)
            conv22 = sg.GATConv_SGRACE(hidden, hidden, 1)
            reluh = sg.Relu_SGRACE()
            lin = torch.nn.Linear(hidden, classes)
            if qbits == 8 and gat == 0:
                for k, v in (("att2.weight", att2.weight), ("att2.attention", att2.attention), ("conv22.weight", conv22.weight),
                             ("conv22.attention", conv22.attention), ("lin.weight", lin.weight), ("lin.bias", lin.bias)):
                    d["p_" + k] = v.detach().numpy().copy()
            with torch.no_grad():
                xt = torch.from_numpy(x)
                edge_index, norm = sg.sym_norm2(torch.from_numpy(ei), n, None, fill, torch.float32)
                a = torch.sparse_coo_tensor(edge_index, norm, (n, n))
                h1 = att2.forward(gat, 0, 1, xt, edge_index, norm, a)
                h1 = reluh(h1)
                h2 = conv22.forward(gat, 1, 0, h1, edge_index, norm, a)
                out = lin(h2.float())                      # eval mode: dropout is the identity
            if qbits == 8:
                d[f"q{qbits}_gat{gat}_h1"] = h1.numpy().astype(np.float32)
            d[f"q{qbits}_gat{gat}_h2"] = h2.numpy().astype(np.float32)
            d[f"q{qbits}_gat{gat}_logits"] = out.numpy().astype(np.float32)
            print(f"demo model q{qbits} gat{gat}: |h2| {float(np.abs(d[f'q{qbits}_gat{gat}_h2']).max()):.4f} "
                  f"|logits| {float(np.abs(d[f'q{qbits}_gat{gat}_logits']).max()):.4f}")
    np.savez_compressed(os.path.join(HERE, "demo_model_cora.npz"), **d)
    print("demo_model_cora.npz")


def real_files():
    if not (O.ref_available("half") and O.ref_available("float")):
        print("oracle/_ref not built; run `make -C oracle ref` first")
        return
    real_mol()
    real_cora()
    real_mutag()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "real":
        real_files()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "demo":
        demo_model()
        sys.exit(0)
    citeseer()
    toy()
    ref_hls()
    backward_emulation()
    real_files()
    demo_model()
    qlayers()
