"""Where does the CUDA path leave the fp64 oracle at BASELINE sizes?  (1) per-layer logit error and label flips of the
whole pipeline for a few scene sizes and both conv algorithms; (2) "teacher forced" op-level check at full size: every
decoder kernel is fed the ORACLE's inputs of that layer (voxel features, queries, labels) and compared with the oracle's
output of the same sub-layer, so a deviation is pinned to one kernel.  Usage: python tools/parity_diag.py [shape]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_forward, oracle_model, rel_err  # noqa: E402
import test_gpu_parity as T  # noqa: E402
from agile3d_b200 import ops  # noqa: E402

DEV = "cuda"
shape = sys.argv[1] if len(sys.argv) > 1 else "headline"
base = T.BASELINE_SHAPES[shape]

# ---- (1) whole pipeline vs size
for target in (10000, 40000, base["target"]):
    cfg = dict(base, target=target)
    sc, coords, clicks, times = T._baseline_scene(cfg)
    ref_m = oracle_model(5, torch.float64)
    pcd_r, _, _, ref_layers = oracle_forward(ref_m, coords, sc["feats"], sc["raw_coords"], [clicks], [times], dtype=torch.float64)
    for algo, name in ((None, "tc"), (ops.ALGO_SIMT, "simt")):
        m = T._gpu_model(5, algo)
        h, layers = T._run_gpu(m, coords, sc["feats"], sc["raw_coords"], [clicks], [times])
        errs = [rel_err(layers[l][0].cpu().numpy(), ref_layers[l][0].numpy()) for l in range(3)]
        flips = [int((layers[l][0].cpu().argmax(1) != ref_layers[l][0].argmax(1)).sum()) for l in range(3)]
        print(f"pipeline {shape} N={coords.shape[0]} conv={name}: pcd {rel_err(h[0].F.cpu().numpy(), pcd_r.F.numpy()):.2e} "
              f"layers {['%.2e' % e for e in errs]} flips {flips}", flush=True)

# ---- (2) teacher-forced decoder ops at full size
sc, coords, clicks, times = T._baseline_scene(base)
ref = oracle_model(5, torch.float64)
m = T._gpu_model(5)
H, d = 8, 128
from oracle import me_ref as ME  # noqa: E402
with torch.no_grad():
    x = ME.SparseTensor(coordinates=torch.as_tensor(coords), features=torch.as_tensor(sc["feats"]).double())
    pcd, aux, (raw, rows), pos_l = ref.forward_backbone(x, torch.as_tensor(sc["raw_coords"]).double())
    ridx = torch.from_numpy(rows[0])
    src, xyz, pos = pcd.F[ridx], raw[ridx], pos_l[0]
    lo, hi = xyz.min(0, keepdim=True)[0], xyz.max(0, keepdim=True)[0]
    ck, ct = clicks, times
    K = len(ck) - 1
    split = [len(ck[str(i)]) for i in range(1, K + 1)]
    fg_rows = [i for o in range(1, K + 1) for i in ck[str(o)]]
    fg_t = [t for o in range(1, K + 1) for t in ct[str(o)]]
    tt = ref.time_encode.double()
    fg_pos = ref.pos_enc(xyz[fg_rows], lo, hi) + tt[fg_t]
    fg_q = src[fg_rows]
    bg_q, bg_pos = ref.bg_query_feat.weight, ref.bg_query_pos.weight
    if len(ck["0"]):
        bg_pos = torch.cat([bg_pos, ref.pos_enc(xyz[ck["0"]], lo, hi) + tt[ct["0"]]], 0)
        bg_q = torch.cat([bg_q, src[ck["0"]]], 0)
    qpos = torch.cat([fg_pos, bg_pos], 0)
    n_fg = fg_q.shape[0]
    nq = qpos.shape[0]
    q_obj = torch.tensor([o for o, n in enumerate(split, start=1) for _ in range(n)] + [0] * (nq - n_fg), dtype=torch.int32)
    g = lambda t: t.float().contiguous().to(DEV)
    pos_g, qpos_g, qobj_g = g(pos), g(qpos).unsqueeze(0), q_obj.to(DEV)
    mask, lab = None, None
    for l in range(3):
        q_in = torch.cat([fg_q, bg_q], 0)
        q1 = ref.c2s_attention[l][0](q_in, src, memory_mask=mask, pos=pos, query_pos=qpos)
        q2 = ref.c2c_attention[l][0](q1, query_pos=qpos)
        q3 = ref.ffn_attention[l][0](q2)
        src_next = ref.s2c_attention[l][0](src, q3, pos=qpos, query_pos=pos)
        fg_q, bg_q = q3[:n_fg], q3[n_fg:]
        logits, mask_next = ref.mask_module(fg_q, bg_q, src_next, split)
        # ---- the same sub-layers on the GPU with the oracle's inputs
        c2s, s2c = m.c2s_attention[l][0], m.s2c_attention[l][0]
        qfold = m._fold_c2s(c2s.multihead_attn, g(q_in).unsqueeze(0), qpos_g, H)
        lab_g = cnt_g = None
        if lab is not None:
            lab_g = lab.to(torch.uint8).to(DEV)
            cnt_g = torch.bincount(lab, minlength=K + 1).to(torch.int32).to(DEV)
        for algo, name in ((ops.ALGO_TC, "tc"), (ops.ALGO_SIMT, "simt")):
            ctx = ops.c2s_attn_fwd(g(src), pos_g, qfold[0], nq, H, lab_g, qobj_g if lab is not None else None, cnt_g, algo=algo)
            q1_g = m._finish_c2s(c2s, g(q_in).unsqueeze(0), ctx.unsqueeze(0), H)[0]
            A, c, U = m._fold_s2c(s2c.multihead_attn, g(q3).unsqueeze(0), qpos_g, H)
            E = m.mask_embed_head(m.decoder_norm(g(q3))).contiguous()
            xo, lg, lb, cnt = ops.s2c_mask_fwd(g(src), pos_g, A[0], c[0], U[0], s2c.multihead_attn.out_proj.bias, s2c.norm.weight,
                                               s2c.norm.bias, s2c.norm.eps, E, qobj_g, nq, H, K + 1, algo=algo)
            flips = int((lb.cpu().long() != logits.argmax(1)).sum())
            print(f"teacher-forced layer {l} {name}: c2s->queries {rel_err(q1_g.cpu().numpy(), q1.numpy()):.2e}  "
                  f"s2c x' {rel_err(xo.cpu().numpy(), src_next.numpy()):.2e}  logits {rel_err(lg.cpu().numpy(), logits.numpy()):.2e}  "
                  f"label flips {flips}  count ok {bool(torch.equal(cnt.cpu(), torch.bincount(lb.cpu().long(), minlength=K + 1).int()))}",
                  flush=True)
        # how sensitive is the oracle's own c2s to the GPU's labels?  (flip study)
        src, mask, lab = src_next, mask_next, logits.argmax(1)
        print(f"   oracle layer {l}: label histogram {torch.bincount(lab, minlength=K + 1).tolist()}", flush=True)
