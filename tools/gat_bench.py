"""Device times of the quantised full-design layer (GAT / GCN, q8 / q4) on a PubMed-shape batch.
Reports the per-stage CUDA-event times the library records, and the algorithmic GB/s."""
import sys, json
sys.path.insert(0, ".")
import numpy as np
from sgracex1_b200 import _lib, graphs as G, quant as Q
from sgracex1_b200.driver import HostLayer
from sgracex1_b200.pynq_compat import MmultTop

copies = int(sys.argv[1]) if len(sys.argv) > 1 else 32
p1 = G.pubmed_shape()
b = G.block_diagonal([p1], copies)
ip = MmultTop(0)
rng = np.random.default_rng(7)
att = rng.uniform(-0.6, 0.6, size=2 * b.P).astype(np.float32)
adj, fea = (b.adj_rowptr, b.adj_col, b.adj_val), (b.fea_rowptr, b.fea_col, b.fea_val)
nnz_a, nnz_f = len(adj[1]), len(fea[1])
for qbits in (8, 4, 0):
    for gat in (1, 0):
        if qbits == 0 and gat == 0:
            continue
        c = Q.layer_constants(qbits) if qbits else None
        ip.configure(qbits=qbits, staging=1)
        hl = HostLayer(ip, _lib.MODE_FULL, N=b.N, M=b.M, P=b.P, nnz_adj=nnz_a, nnz_fea=nnz_f, dense=False, gat=True, coo=True)
        rm = ip.register_map
        if c:
            rm.scale_fea = c["scale_fea"]; rm.deq_factor = Q.float_bits(c["deq_o"])
            rm.quantization_scale_fea = Q.float_bits(1 / c["f_s"]); rm.quantization_scale_w = Q.float_bits(1 / c["w_s"])
            rm.quantization_scale_adj = Q.float_bits(1 / c["a_s"]); rm.quantized_multiplier = c["internal_quantization"]
        hl.load(N=b.N, M=b.M, P=b.P, adj=adj, B=b.B, fea=fea, relu=1, attention=att, gat_mode=gat)
        best = None
        for _ in range(4):
            hl.run()
            t = ip.handle.stage_times()
            best = t if best is None or t[2] < best[2] else best
        fea_b = (b.N + 1) * 4 + nnz_f * 8 + b.M * b.P * 4 + b.N * b.P * 4
        adj_b = (b.N + 1) * 4 + nnz_a * 8 + 2 * b.N * b.P * 4 + (2 * nnz_a * 4 + 2 * b.N * 4 if gat else 0)
        print(json.dumps({"q": qbits, "gat": gat, "N": b.N, "nnz_fea": nnz_f, "nnz_adj": nnz_a, "fea_ms": round(best[0], 4),
                          "adj_ms": round(best[1], 4), "total_ms": round(best[2], 4), "fea_GBs": round(fea_b / best[0] / 1e6, 1),
                          "adj_GBs": round(adj_b / best[1] / 1e6, 1), "gteps": round(nnz_a / best[1] / 1e6, 2)}), flush=True)
        hl.free()
