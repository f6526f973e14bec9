"""In-memory editing pipeline: the BASELINE config-1 "edit" as one call.

    images = EditPipeline(unet).edit(x0, mask)

runs, for one image, exactly the sequence of the reference's `run_edit_null_space_projection`
(src/modules/edit.py:2216-2366) for one principal direction:
  DDIM inversion (98 U-Net calls) -> forward to t = edit_t (40 calls) -> rank-k local basis of the
  masked PMP Jacobian (n_iter power iterations) -> rank-k_null basis of the complement mask ->
  null-space projection -> +/- edits along direction `pc` -> 59 DDIM steps on the batch of 2*vis_num-1
  latents with eta = 1 for t <= 0.2 T.
Nothing is written to disk here (the file-producing driver is `EditUncondDiffusion`); host tensors
go in and come out, so the call is the unit bench.py times end to end.
"""
import os
import types

import torch

from . import ops
from .edit import EditUncondDiffusion, local_basis, local_basis_pair


class EditPipeline(object):
    def __init__(self, unet, k=5, k_null=5, edit_t=0.6, n_iter=12, scale=0.5, num_step=16, vis_num=2,
                 for_steps=100, inv_steps=100, boost_t=0.2, result_folder="/tmp/loco_b200_runs", basis_streams=None):
        self.unet = unet
        # concurrent per-image power methods in edit_batch (each stream pins one more (1, k + k_null, k + k_null) plan)
        self.basis_streams = int(os.environ.get("LOCO_BASIS_STREAMS", "2")) if basis_streams is None else int(basis_streams)
        self._streams = []
        self.device = unet.device
        self.k, self.k_null, self.n_iter = k, k_null, n_iter
        self.vis_num = vis_num
        args = types.SimpleNamespace(
            device=self.device, dtype=torch.float32, seed=1, model_name="CelebA_HQ_HF",
            dataset_name="CelebA_HQ_mask", image_size=unet.arch["resolution"], for_steps=for_steps,
            inv_steps=inv_steps, edit_t=edit_t, performance_boosting_t=boost_t,
            x_space_guidance_edit_step=1.0, x_space_guidance_scale=scale,
            x_space_guidance_num_step=num_step, result_folder=result_folder, sample_idx=0,
            verbose=False, save_images=False, noise_schedule=None)
        self.driver = EditUncondDiffusion(args, unet=unet)
        self.R = unet.arch["resolution"]

    def _v0(self, k, gen):
        d = 3 * self.R * self.R
        # Algorithm-1 init (reference modules/edit.py:2435-2437): orthonormalised Gaussian block
        g = torch.randn(d, k, device=self.device, dtype=torch.float32, generator=gen)
        V, _ = ops.orthonormalise(g.T.contiguous())
        return V

    @torch.no_grad()
    def edit_device(self, x0, mask, pc=0, gen=None, v0_mod=None, v0_null=None, noises=None):
        """x0 [1,3,R,R] fp32 and mask bool [3,R,R], both already on the device."""
        drv = self.driver
        # inversion: reference run_DDIMinversion (:2117-2167) with the image injected
        sched = drv.scheduler
        sched.set_timesteps(drv.inv_steps, device=self.device, is_inversion=True)
        xt = x0.contiguous()
        n = len(sched._ts_host)
        for i in range(n - 1):
            t = sched._ts_host[i]
            xt = sched.step(self.unet(xt, t), t, xt, eta=0, t_idx=i).prev_sample
        xt, t, t_idx = drv.DDIMforwardsteps(xt, t_start_idx=0, t_end_idx=drv.edit_t_idx, save_image=False)
        t_host = sched._ts_host[t_idx]
        vT_mod, s_mod, vT_null, s_null = local_basis_pair(
            self.unet, sched, xt, t_host, self.k, self.k_null, mask,
            v0=v0_mod if v0_mod is not None else self._v0(self.k, gen),
            v0_null=v0_null if v0_null is not None else self._v0(self.k_null, gen), n_iter=self.n_iter)
        vT = ops.nullspace_project(vT_mod, vT_null, project=True)
        batch = drv.build_edit_batch(xt, vT[pc], self.vis_num)
        if noises is not None:
            drv.noise_fn = lambda i, x: noises[i]
        elif gen is not None:
            drv.noise_fn = lambda i, x: torch.randn(x.shape, device=x.device, dtype=x.dtype, generator=gen)
        else:
            drv.noise_fn = None
        imgs = drv.DDIMforwardsteps(batch, t_start_idx=drv.edit_t_idx, t_end_idx=-1, save_image=False,
                                    performance_boosting=True)
        return dict(images=imgs, vT=vT, vT_modify=vT_mod, vT_null=vT_null, s_modify=s_mod, s_null=s_null,
                    xt=xt)

    @torch.no_grad()
    def edit_sharded_device(self, x0, mask, pc=0, gen=None, v0_mod=None, v0_null=None, group=None, timer=None):
        """ONE image edited by all ranks of `group` together (single-edit latency lever):
        the inversion / forward-to-t chain is serial and runs replicated on every rank (same kernels,
        same inputs); the k + k_null probes of the two local bases are sharded jointly over the ranks
        with one all-gather of the W rows per power iteration; the 2*vis_num-1 edited latents of the
        final DDIM stage are sharded over the ranks and gathered once at the end.
        Every rank returns the full image batch."""
        import torch.distributed as dist
        from . import dist as ld
        drv = self.driver
        sched = drv.scheduler
        sched.set_timesteps(drv.inv_steps, device=self.device, is_inversion=True)
        xt = x0.contiguous()
        n = len(sched._ts_host)
        for i in range(n - 1):
            t = sched._ts_host[i]
            xt = sched.step(self.unet(xt, t), t, xt, eta=0, t_idx=i).prev_sample
        xt, t, t_idx = drv.DDIMforwardsteps(xt, t_start_idx=0, t_end_idx=drv.edit_t_idx, save_image=False)
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            # replicated chains agree to the TF32 noise level only (GroupNorm statistics use atomics);
            # rank 0's x_t is THE x_t, so that all ranks probe the same Jacobian
            dist.broadcast(xt, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        t_host = sched._ts_host[t_idx]
        v0_mod = v0_mod if v0_mod is not None else self._v0(self.k, gen)
        v0_null = v0_null if v0_null is not None else self._v0(self.k_null, gen)
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            src = dist.get_global_rank(group, 0) if group is not None else 0
            dist.broadcast(v0_mod, src=src, group=group)
            dist.broadcast(v0_null, src=src, group=group)
        vT_mod, s_mod, vT_null, s_null = ld.sharded_local_basis_pair_cuda(
            self.unet, sched, xt, t_host, self.k, self.k_null, mask, v0_mod, v0_null, self.n_iter,
            group=group, timer=timer)
        vT = ops.nullspace_project(vT_mod, vT_null, project=True)
        batch = drv.build_edit_batch(xt, vT[pc], self.vis_num)
        mine = ld.shard_images(list(range(batch.shape[0])), group)
        drv.noise_fn = (lambda i, x: torch.randn(x.shape, device=x.device, dtype=x.dtype, generator=gen)) \
            if gen is not None else None
        if mine:
            local = drv.DDIMforwardsteps(batch[mine[0]:mine[-1] + 1].contiguous(), t_start_idx=drv.edit_t_idx,
                                         t_end_idx=-1, save_image=False, performance_boosting=True)
        else:
            local = batch[:0]
        imgs = ld.gather_images(local, batch.shape[0], group)
        return dict(images=imgs, vT=vT, vT_modify=vT_mod, vT_null=vT_null, s_modify=s_mod, s_null=s_null, xt=xt)

    @torch.no_grad()
    def edit_sharded(self, x0_host, mask_host, pc=0, gen=None, group=None, timer=None):
        """Host tensors in, edited images back on the host; all ranks of `group` call this together."""
        x0 = x0_host.to(self.device, non_blocking=True)
        mask = mask_host.to(self.device, non_blocking=True)
        out = self.edit_sharded_device(x0, mask, pc=pc, gen=gen, group=group, timer=timer)
        return out["images"].to("cpu", non_blocking=False)

    @torch.no_grad()
    def edit_batch_device(self, x0s, masks, pc=0, gen=None):
        """Batch editing (BASELINE config 2: many image/mask pairs per GPU).  x0s [B,3,R,R], masks
        bool [B,3,R,R] on the device.  The DDIM inversion, the forward pass to t and the final
        denoising of all 5*B edited latents run as single batched U-Net programs; the two local
        bases are per image (their tangents belong to one primal point).  Returns [B,5,3,R,R]."""
        drv = self.driver
        sched = drv.scheduler
        B = x0s.shape[0]
        sched.set_timesteps(drv.inv_steps, device=self.device, is_inversion=True)
        xt = x0s.contiguous()
        n = len(sched._ts_host)
        for i in range(n - 1):
            t = sched._ts_host[i]
            xt = sched.step(self.unet(xt, t), t, xt, eta=0, t_idx=i).prev_sample
        xt, t, t_idx = drv.DDIMforwardsteps(xt, t_start_idx=0, t_end_idx=drv.edit_t_idx, save_image=False)
        t_host = sched._ts_host[t_idx]
        batches, vTs = [], []
        # the local bases of different images are independent: image b runs on stream b % basis_streams with its own
        # plan slot (saved activations, split-K scratch and iteration buffers), so that the <= 32^2 layers of one
        # image's Jacobian passes (which cannot fill 148 SMs with 11 rows) overlap with the other image's kernels
        ns = max(1, min(self.basis_streams, B))
        main = torch.cuda.current_stream(self.device)
        if ns > 1:
            if len(self._streams) < ns:
                self._streams += [torch.cuda.Stream(self.device) for _ in range(ns - len(self._streams))]
            for s_ in self._streams[:ns]:
                s_.wait_stream(main)
        for b in range(B):
            slot = b % ns
            with torch.cuda.stream(self._streams[slot] if ns > 1 else main):
                xb = xt[b:b + 1].contiguous()
                vT_mod, _, vT_null, _ = local_basis_pair(
                    self.unet, sched, xb, t_host, self.k, self.k_null, masks[b], v0=self._v0(self.k, gen),
                    v0_null=self._v0(self.k_null, gen), n_iter=self.n_iter, slot=slot)
                vT = ops.nullspace_project(vT_mod, vT_null, project=True)
                vTs.append(vT)
                batches.append(drv.build_edit_batch(xb, vT[pc], self.vis_num))
        if ns > 1:
            for s_ in self._streams[:ns]:
                main.wait_stream(s_)
            for t_ in batches + vTs:
                t_.record_stream(main)
        per = batches[0].shape[0]
        allx = torch.cat(batches, 0).contiguous()
        drv.noise_fn = (lambda i, x: torch.randn(x.shape, device=x.device, dtype=x.dtype, generator=gen)) \
            if gen is not None else None
        imgs = drv.DDIMforwardsteps(allx, t_start_idx=drv.edit_t_idx, t_end_idx=-1, save_image=False,
                                    performance_boosting=True)
        return dict(images=imgs.reshape(B, per, *imgs.shape[1:]), vT=torch.stack(vTs, 0), xt=xt)

    @torch.no_grad()
    def edit_batch(self, x0s_host, masks_host, pc=0, gen=None):
        """Host (pinned) tensors in, edited images [B,5,3,R,R] back on the host."""
        x0s = x0s_host.to(self.device, non_blocking=True)
        masks = masks_host.to(self.device, non_blocking=True)
        out = self.edit_batch_device(x0s, masks, pc=pc, gen=gen)
        return out["images"].to("cpu", non_blocking=False)

    @torch.no_grad()
    def edit(self, x0_host, mask_host, pc=0, gen=None, **kw):
        """Host tensors in (pinned for async copies), edited images [2*vis_num-1,3,R,R] back on the host."""
        x0 = x0_host.to(self.device, non_blocking=True)
        mask = mask_host.to(self.device, non_blocking=True)
        out = self.edit_device(x0, mask, pc=pc, gen=gen, **kw)
        imgs = out["images"].to("cpu", non_blocking=False)
        return imgs
