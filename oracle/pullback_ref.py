"""ORACLE (test infrastructure, never on the product path).

CPU restatement of the reference's editing-direction algorithms around the U-Net: scheduler, PMP,
power-method local basis, null-space projection, DDIM loops, edit application.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
Pinned against the unmodified reference classes by tests/golden/make_golden.py.

Citations are relative to /root/reference/src.
"""
import torch


class RefScheduler:
    """utils/utils.py:305-423 (YHCustomScheduler), linear schedule, learn_sigma = False."""

    def __init__(self, dtype=torch.float32):
        self.t_max = 999
        betas = torch.linspace(0.0001, 0.02, 1000, dtype=torch.float64)      # :405-406, :388-392
        self.betas = betas.to(dtype)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0).to(dtype)     # :401-403
        self.timesteps = None
        self.timesteps_next = None

    def set_timesteps(self, n, is_inversion=False):                           # :316-329
        seq = torch.linspace(0, 1, n) * self.t_max
        if is_inversion:
            seq = seq + 1e-6
            seq_prev = torch.cat([torch.tensor([-1.0]), seq[:-1]], dim=0)
            self.timesteps = seq_prev[1:]
            self.timesteps_next = seq[1:]
        else:
            seq_prev = torch.cat([torch.tensor([-1.0]), seq[:-1]], dim=0)
            self.timesteps = torch.flip(seq[1:], dims=[0])
            self.timesteps_next = torch.flip(seq_prev[1:], dims=[0])

    def alpha(self, t):
        """utils/utils.py:444-461 (extract): gather at t.long()."""
        return self.alphas_cumprod[int(torch.as_tensor(t).long())]

    def step(self, et, t, xt, eta=0.0, noise=None):                           # :342-374
        idx = self.timesteps.tolist().index(float(t))
        t_next = self.timesteps_next[idx]
        at, at_next = self.alpha(t), self.alpha(t_next)
        P_xt = (xt - et * (1 - at).sqrt()) / at.sqrt()
        if eta == 0:
            D_xt = (1 - at_next).sqrt() * et
            xt_next = at_next.sqrt() * P_xt + D_xt
        else:
            sigma_t = ((1 - at / at_next) * (1 - at_next) / (1 - at)).sqrt()
            D_xt = (1 - at_next - eta * sigma_t ** 2).sqrt() * et
            if noise is None:
                noise = torch.randn_like(xt)
            xt_next = at_next.sqrt() * P_xt + D_xt + eta * sigma_t * noise
        return xt_next, P_xt


def get_x0(unet, sched, t, x, mask=None, noise=False):
    """modules/edit.py:2369-2403 (get_x0 / get_et): PMP, then row-major boolean selection."""
    et = unet(x, t)
    if noise:
        out = et
    else:
        at = sched.alpha(t)
        out = (x - et * (1 - at).sqrt()) / at.sqrt()
    if mask is not None:
        out = out[:, mask]
    return out


def power_iteration(unet, sched, x, t, V, mask=None, noise=False):
    """One pass of the loop body of local_encoder_decoder_pullback_xt (modules/edit.py:2443-2483).

    V: [k, d] orthonormal rows.  Returns (u [k, l_o] = J V^T rows, w [k, d] = U^T J,
    s = svdvals(w), Vh [k, d])."""
    k = V.shape[0]
    shp = x.shape[1:]
    us = []
    for j in range(k):                                   # jacfwd over a: u_j = J v_j  (:2449-2456)
        vj = V[j].reshape(1, *shp)
        _, u = torch.func.jvp(lambda xx: get_x0(unet, sched, t, xx, mask, noise), (x,), (vj,))
        us.append(u.reshape(-1))
    u = torch.stack(us, 0).detach()
    xg = x.detach().clone().requires_grad_(True)         # jacobian of <u_b, f(x)>   (:2460-2480)
    out = get_x0(unet, sched, t, xg, mask, noise).reshape(-1)
    ws = []
    for j in range(k):
        (g,) = torch.autograd.grad((u[j] * out).sum(), xg, retain_graph=j + 1 < k)
        ws.append(g.reshape(-1))
    w = torch.stack(ws, 0)
    _, s, vh = torch.linalg.svd(w, full_matrices=False)  # :2482
    return u, w, s, vh


def local_basis(unet, sched, x, t, V0, n_iter, mask=None, noise=False):
    """modules/edit.py:2406-2504 with an injected V0 and a fixed iteration count.
    Returns (u.T [l_o,k], s.sqrt() [k], vT [k,d]) like the reference (:2499-2502)."""
    V = V0
    for _ in range(n_iter):
        u, w, s, V = power_iteration(unet, sched, x, t, V, mask, noise)
    return u.T, s.sqrt(), V


def nullspace_project(vT_mod, vT_null, k_null, project=True):
    """modules/edit.py:2317-2323."""
    if not project:
        return vT_mod / vT_mod.norm(dim=1, keepdim=True)
    vn = vT_null[:k_null, :]
    vT = (vn.T @ (vn @ vT_mod.T)).T
    vT = vT_mod - vT
    return vT / vT.norm(dim=1, keepdim=True)


def ddim_inversion(unet, sched, x0, steps=100):
    """modules/edit.py:2117-2167 (run_DDIMinversion): loop breaks before the last timestep."""
    sched.set_timesteps(steps, is_inversion=True)
    ts = sched.timesteps
    xt = x0
    with torch.no_grad():
        for i, t in enumerate(ts):
            if i == len(ts) - 1:
                break
            xt, _ = sched.step(unet(xt, t), t, xt, eta=0)
    return xt


def ddim_forward(unet, sched, xt, t_start_idx, t_end_idx, steps=100, boost_idx=None, noises=None):
    """modules/edit.py:2508-2614 (DDIMforwardsteps): eta = 1 for i >= boost_idx
    (performance_boosting, :2556-2559); `noises[i]` injects the eta = 1 noise for parity."""
    sched.set_timesteps(steps)
    ts = sched.timesteps
    with torch.no_grad():
        for i, t in enumerate(ts):
            if t_end_idx == i:
                return xt, t, i
            if i < t_start_idx:
                continue
            eta = 1 if (boost_idx is not None and boost_idx <= i and boost_idx != len(ts) - 1) else 0
            nz = noises[i] if (noises is not None and eta) else None
            xt, _ = sched.step(unet(xt, t), t, xt, eta=eta, noise=nz)
    return xt


def edit_batch(xt, vT_row, scale, num_step, vis_num, edit_step=1.0):
    """modules/edit.py:2340-2363: build the batch of edited latents for one direction."""
    xts = {}
    for direction in (1, -1):
        vk = direction * vT_row.reshape(1, *xt.shape[1:])
        lst = [xt.clone()]
        for _ in range(num_step):
            lst.append(lst[-1] + scale * edit_step * vk)      # :2618-2625
        x = torch.cat(lst, 0)
        x = x[[0, -1], :] if vis_num == 1 else x[::(x.size(0) // vis_num)]
        xts[direction] = x
    return torch.cat([xts[-1].flip(dims=[0])[:-1], xts[1]], dim=0)
