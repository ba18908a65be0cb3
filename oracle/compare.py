"""Test infrastructure (like everything under oracle/): comparison of an implementation's mask logits with the CPU
oracle across the decoder's discrete decisions.  Imported by tests/, __graft_entry__.smoke() and bench.py's parity
field only - never by the product path."""
import numpy as np
import torch


def rel_err(a, b):
    """max |a-b| / max |b|: the 'relative' in "within 1e-3 relative on fp32 mask logits"."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _forward(model, bb, clicks, times, force_labels=None):
    with torch.no_grad():
        kw = {} if force_labels is None else {"force_labels": force_labels}
        out = model.forward_mask(*bb, clicks, times, **kw)
    return [a["pred_masks"] for a in out["aux_outputs"]] + [out["pred_masks"]]


def decision_forced_errors(model, coords, feats, raw, clicks, times, got_layers, dtype=torch.float64, margin_tol=1e-3):
    """Parity of a decoder with DISCRETE decisions between its layers (agile3d.py:365-380: argmax labels of layer l
    mask the attention of layer l+1).  A voxel whose two best logits are within rounding of each other may
    legitimately land on either side, and when an object's predicted mask holds a few dozen voxels one such voxel moves
    that object's query by percents - in the reference as much as here.  So:
      free[l]    rel. error of the implementation's logits against the free-running oracle (informative)
      forced[l]  rel. error against the oracle run with the IMPLEMENTATION's label decisions: the 1e-3 criterion
      bad_flips  differing decisions on voxels whose oracle top-2 margin exceeds margin_tol * max|logit|: must be 0
    got_layers: [layer][scene] logits of the implementation under test (torch tensors, any device)."""
    got = [[t.detach().cpu() for t in layer] for layer in got_layers]
    from oracle import me_ref as ME

    x = ME.SparseTensor(coordinates=torch.as_tensor(coords), features=torch.as_tensor(feats).to(dtype))
    with torch.no_grad():
        bb = model.forward_backbone(x, torch.as_tensor(raw).to(dtype))
    pcd = bb[0]
    free = _forward(model, bb, clicks, times)
    labels = [[g.argmax(1) for g in layer] for layer in got]
    forced = _forward(model, bb, clicks, times, labels)
    res = {"free": [], "forced": [], "flips": [], "bad_flips": 0, "pcd": pcd}
    for l in range(len(got)):
        fe = fo = 0.0
        nf = 0
        for b in range(len(got[l])):
            fe = max(fe, rel_err(got[l][b].numpy(), free[l][b].numpy()))
            fo = max(fo, rel_err(got[l][b].numpy(), forced[l][b].numpy()))
            ref = forced[l][b]                      # decisions are judged where both runs saw the same inputs
            diff = got[l][b].argmax(1) != ref.argmax(1)
            nf += int(diff.sum())
            if ref.shape[1] > 1 and bool(diff.any()):
                top2 = torch.topk(ref[diff], 2, dim=1)[0]
                res["bad_flips"] += int(((top2[:, 0] - top2[:, 1]) > margin_tol * float(ref.abs().max())).sum())
        res["free"].append(fe)
        res["forced"].append(fo)
        res["flips"].append(nf)
    return res
