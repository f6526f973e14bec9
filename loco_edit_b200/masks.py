"""Mask files and mask post-processing of the editing-direction path (host logic, no kernels).

The reference obtains its masks three ways (src/modules/edit.py:2234-2267, 1395-1407):
  * CelebA_HQ_mask: ground-truth semantic masks of the dataset (`dataset.getmask`);
  * every other dataset: `mask/mask.pt` under the result folder, written by the SAM wrapper
    (src/modules/mask_segmentation.py:18-26) as a bool tensor [n_masks, res, res]; row `--mask_index`
    is repeated over the 3 channels (src/modules/edit.py:2247, 2263);
  * T-LOCO / DeepFloyd: the DiffEdit mask, the thresholded difference of two guided noise
    predictions (src/modules/edit.py:1395-1407).
The SAM network itself is out of scope here; this module reads and writes the
same file format, reproduces the resize / rounding the reference applies to SAM's output, and
restates the DiffEdit formula so that a mask produced by either route can be fed to the drop-in
driver (`EditUncondDiffusion.run_edit_null_space_projection`).
"""
import os

import torch


def resize_masks(masks, resolution):
    """SAM post-processing (mask_segmentation.py:23-24): nearest-neighbour resize of bool/float masks
    [n, H, W] to [n, resolution, resolution], rounded, bool.  (`F.interpolate` default mode 'nearest':
    source index floor(dst * in / out).)"""
    m = torch.as_tensor(masks)
    n, H, W = m.shape
    ys = torch.div(torch.arange(resolution) * H, resolution, rounding_mode="floor")
    xs = torch.div(torch.arange(resolution) * W, resolution, rounding_mode="floor")
    out = m.to(torch.float32)[:, ys][:, :, xs]
    return torch.round(out).to(torch.bool)


def save_masks(result_folder, masks):
    """Write `mask/mask.pt` (bool [n, res, res]) where the drivers look for it
    (src/modules/edit.py:2238, 2254; file written at mask_segmentation.py:25)."""
    m = torch.as_tensor(masks)
    if m.dim() == 4 and m.shape[1] == 1:
        m = m[:, 0]
    if m.dim() != 3:
        raise ValueError("masks must be [n, res, res] (or [n, 1, res, res]), got %s" % (tuple(m.shape),))
    d = os.path.join(result_folder, "mask")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "mask.pt")
    torch.save(m.to(torch.bool).cpu(), path)
    return path


def load_mask(result_folder, mask_index=0):
    """Row `mask_index` of mask/mask.pt as the [3, res, res] bool mask the pull-back takes
    (src/modules/edit.py:2247: `masks[idx].squeeze(dim=0).repeat(3, 1, 1)`)."""
    masks = torch.load(os.path.join(result_folder, "mask/mask.pt"), map_location="cpu")
    return masks[mask_index].squeeze(dim=0).repeat(3, 1, 1)


def diffedit_mask(eps_for, eps_edit):
    """DiffEdit mask from two guided noise predictions [n, C, H, W] of n noised copies of the image
    (src/modules/edit.py:1401-1402), INCLUDING the reference's operator precedence: it subtracts
    min / (max - min) rather than normalising to [0, 1] before rounding.  Returns bool [1, H, W]."""
    mask = (eps_for - eps_edit).mean(dim=0, keepdim=True).mean(dim=1)
    return torch.round(mask - mask.min() / (mask.max() - mask.min())).to(torch.bool)


def rectangle_mask(resolution, rows=(3, 5), cols=(2, 6), eighths=True):
    """Synthetic stand-in used by the benches and tests: True on rows 3/8..5/8 and
    columns 2/8..6/8 of a [3, res, res] image."""
    R = resolution
    m = torch.zeros(3, R, R, dtype=torch.bool)
    m[:, rows[0] * R // 8:rows[1] * R // 8, cols[0] * R // 8:cols[1] * R // 8] = True
    return m
