"""Result files in the reference's formats, so `helpers/visualizations.py` of the reference can consume a run of
this package unchanged: `original.pt`, `measurement.pt`, `reconstructions.pt`, `ZF.pt` (torch.save of CPU tensors)
and `args_dict.pkl` (reference: scripts/acdc_SENSE_real_img.py:126-127,162-175;
scripts/cine_SENSE_real_img_2d_time.py writes the same set)."""
import os
import pickle

import torch


def save_reconstruction(save_dir, original, measurement, reconstructions, linear_tfm=None, zero_filled=None, args_dict=None):
    """original (1|B,C,H,W) complex, measurement (Nc,B,C,H,W), reconstructions (B,C,H,W) complex.  `ZF.pt` is
    `linear_tfm.conj_op(measurement)[0]` as in the reference unless `zero_filled` is given."""
    os.makedirs(save_dir, exist_ok=True)
    if zero_filled is None:
        if linear_tfm is None:
            raise ValueError("save_reconstruction needs linear_tfm or zero_filled for ZF.pt")
        zero_filled = linear_tfm.conj_op(measurement)[0]
    cpu = lambda t: torch.as_tensor(t).detach().cpu()
    torch.save(cpu(original), os.path.join(save_dir, "original.pt"))
    torch.save(cpu(measurement), os.path.join(save_dir, "measurement.pt"))
    torch.save(cpu(reconstructions), os.path.join(save_dir, "reconstructions.pt"))
    torch.save(cpu(zero_filled), os.path.join(save_dir, "ZF.pt"))
    with open(os.path.join(save_dir, "args_dict.pkl"), "wb") as wf:
        pickle.dump(dict(args_dict or {}), wf)


def load_reconstruction(save_dir):
    out = {k: torch.load(os.path.join(save_dir, k + ".pt")) for k in ("original", "measurement", "reconstructions", "ZF")}
    with open(os.path.join(save_dir, "args_dict.pkl"), "rb") as rf:
        out["args_dict"] = pickle.load(rf)
    return out
