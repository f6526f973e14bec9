"""Mirror of the reference's `YHCustomScheduler` (src/utils/utils.py:305-423) with the update
arithmetic running in the `loco_ddim_step` CUDA kernel.

Same protocol: `.set_timesteps(n, device=None, is_inversion=False)`, `.timesteps`,
`.timesteps_next`, `.step(et, t, xt, eta=0.0, **kw) -> obj(.prev_sample, .x0)`, `.get_timesteps(t)`,
`.return_alphas_cumprod()`.  Differences (host-sync removal only, same numbers):
  * a host copy of the timestep list is kept, so `step` finds its index without `.tolist()` on a
    device tensor; callers that know the index may pass `t_idx=` and skip the lookup entirely;
  * alpha_bar values are read from a host copy of the fp32 table (the reference gathers them on the
    device with `extract`, utils.py:444-461).
"""
import torch

from . import ops


class SchedulerOutput(object):
    """src/utils/utils.py:300-303."""

    def __init__(self, xt_next, P_xt):
        self.prev_sample = xt_next
        self.x0 = P_xt


def cosine_betas(n=1000, max_beta=0.999):
    """"squaredcos_cap_v2" (Nichol & Dhariwal cosine schedule, the beta schedule of the DeepFloyd-IF
    stage-I DDPM scheduler whose `alphas_cumprod` the reference keeps, utils.py:163):
    beta_i = min(1 - ab((i+1)/n) / ab(i/n), max_beta), ab(s) = cos((s + 0.008) / 1.008 * pi / 2)^2."""
    import math
    ab = lambda s: math.cos((s + 0.008) / 1.008 * math.pi / 2) ** 2
    return torch.tensor([min(1 - ab((i + 1) / n) / ab(i / n), max_beta) for i in range(n)], dtype=torch.float64)


def scaled_linear_betas(n=1000, beta_start=0.00085, beta_end=0.012):
    """"scaled_linear": the beta schedule of the Stable Diffusion 1.x scheduler config whose `alphas_cumprod`
    the reference keeps (get_stable_diffusion_scheduler, utils.py:147-157); diffusers builds it in fp32:
    linspace(sqrt(beta_start), sqrt(beta_end), n) ** 2, alphas_cumprod = cumprod(1 - betas)."""
    return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2


class YHCustomScheduler(object):
    def __init__(self, args=None, device=None, dtype=torch.float32):
        # t_max: 999 for the custom scheduler (utils.py:309); the DeepFloyd-IF monkey patch uses 990
        # (get_deepfloyd_if_scheduler, utils.py:159-170)
        self.t_max = getattr(args, "t_max", 999) if args is not None else 999
        ns = getattr(args, "noise_schedule", None) if args is not None else None
        self.noise_schedule = "linear" if ns is None else ns
        if self.noise_schedule not in ("linear", "squaredcos_cap_v2", "scaled_linear"):
            raise NotImplementedError("noise schedule '%s'" % self.noise_schedule)
        self.device = torch.device(device if device is not None else getattr(args, "device", "cuda:0"))
        self.dtype = getattr(args, "dtype", dtype) if args is not None else dtype
        self.timesteps = None
        self.timesteps_next = None
        self.learn_sigma = False
        # utils.py:385-406: fp64 linspace betas -> cumprod -> cast to args.dtype
        if self.noise_schedule == "linear":
            betas = torch.linspace(0.0001, 0.02, 1000, dtype=torch.float64)
        elif self.noise_schedule == "scaled_linear":
            betas = scaled_linear_betas(1000)
        else:
            betas = cosine_betas(1000)
        self.betas = betas.to(device=self.device, dtype=self.dtype)
        acp = torch.cumprod(1.0 - betas, dim=0).to(self.dtype)
        self._acp_host = acp.float().tolist()
        self.alphas_cumprod = acp.to(self.device)

    def set_timesteps(self, num_inferences, device=None, is_inversion=False):
        # utils.py:316-329; built on the host in fp32, then moved (values identical to torch CPU)
        seq = torch.linspace(0, 1, num_inferences) * self.t_max
        if is_inversion:
            seq = seq + 1e-6
            seq_prev = torch.cat([torch.tensor([-1.0]), seq[:-1]], dim=0)
            ts, tn = seq_prev[1:], seq[1:]
        else:
            seq_prev = torch.cat([torch.tensor([-1.0]), seq[:-1]], dim=0)
            ts, tn = torch.flip(seq[1:], dims=[0]), torch.flip(seq_prev[1:], dims=[0])
        self._ts_host, self._tn_host = ts.tolist(), tn.tolist()
        dev = self.device if device is None else device
        self.timesteps = ts.to(dev)
        self.timesteps_next = tn.to(dev)

    def get_timesteps(self, t):
        # utils.py:331-337
        t_idx = torch.where(self.timesteps == t)
        return self.timesteps_next[t_idx]

    def return_alphas_cumprod(self):
        return self.alphas_cumprod

    def alpha_at(self, t):
        """extract(alphas_cumprod, t, .) for a scalar t: table entry at t.long() (utils.py:458)."""
        return self._acp_host[int(float(t))]

    def index_of(self, t):
        return self._ts_host.index(float(t))      # exact float equality, like utils.py:353

    def step(self, et, t, xt, eta=0.0, t_idx=None, noise=None, **kwargs):
        assert et.shape == xt.shape, "et, xt shape should be same"
        if t_idx is None:
            t_idx = self.index_of(t)
        at = self.alpha_at(self._ts_host[t_idx])
        at_next = self.alpha_at(self._tn_host[t_idx])
        if eta != 0 and noise is None:
            noise = torch.randn_like(xt)                      # utils.py:374
        xn, x0 = ops.ddim_step(xt.contiguous(), et.contiguous(), at, at_next, float(eta),
                               noise=noise, want_x0=True)
        return SchedulerOutput(xn, x0)


class LCMScheduler(object):
    """Restatement of diffusers' `LCMScheduler` as the latent-consistency class uses it
    (`self.scheduler.step(model_pred, t, latents, return_dict=False) -> (prev_sample, denoised)`,
    src/modules/edit.py:140, 237, 194; that class only works with the STOCK scheduler: the custom `step` of
    utils.py:186-213 returns an object that cannot be unpacked).  diffusers is a third-party dependency that is
    not under /root/reference: PARITY UNPINNED for this arithmetic; what is restated is the published algorithm
    (Luo et al., Latent Consistency Models; diffusers `scheduling_lcm.py`, epsilon prediction):
      timesteps: every (original_steps // n)-th of the `original_steps` training-grid points k * (T // original_steps) - 1,
                 descending;
      boundary scalings at s = timestep_scaling * t: c_skip = sd^2 / (s^2 + sd^2), c_out = s / sqrt(s^2 + sd^2), sd = 0.5;
      denoised = c_out * (x - sqrt(1 - a_t) eps) / sqrt(a_t) + c_skip * x;
      prev_sample = sqrt(a_prev) * denoised + sqrt(1 - a_prev) * noise   (the last step returns `denoised`).
    alpha_bar: Stable Diffusion's "scaled_linear" table."""

    def __init__(self, device, original_inference_steps=50, timestep_scaling=10.0, sigma_data=0.5, num_train_timesteps=1000):
        self.device = torch.device(device)
        self.original_inference_steps = original_inference_steps
        self.timestep_scaling, self.sigma_data, self.T = timestep_scaling, sigma_data, num_train_timesteps
        betas = scaled_linear_betas(num_train_timesteps)
        acp = torch.cumprod(1.0 - betas, dim=0)
        self._acp_host = acp.tolist()
        self.alphas_cumprod = acp.to(self.device)
        self.timesteps = None

    def set_timesteps(self, num_inference_steps, device=None):
        k = self.T // self.original_inference_steps
        origin = [i * k - 1 for i in range(1, self.original_inference_steps + 1)]
        skip = self.original_inference_steps // num_inference_steps
        ts = origin[::-skip][:num_inference_steps]
        self._ts_host = [int(t) for t in ts]
        self.num_inference_steps = num_inference_steps
        self.timesteps = torch.tensor(self._ts_host, dtype=torch.long, device=self.device if device is None else device)

    def alpha_at(self, t):
        return self._acp_host[int(t)]

    def scalings(self, t):
        """(c_skip, c_out) of the boundary condition at timestep t."""
        s = self.timestep_scaling * float(int(t))
        sd = self.sigma_data
        return sd * sd / (s * s + sd * sd), s / (s * s + sd * sd) ** 0.5

    def denoised_coefficients(self, t):
        """denoised = c1 * x + c2 * eps."""
        a = self.alpha_at(t)
        c_skip, c_out = self.scalings(t)
        return c_out / a ** 0.5 + c_skip, -c_out * (1.0 - a) ** 0.5 / a ** 0.5

    def step(self, model_output, timestep, sample, t_idx=None, noise=None, return_dict=False, **kwargs):
        if t_idx is None:
            t_idx = self._ts_host.index(int(timestep))
        c1, c2 = self.denoised_coefficients(self._ts_host[t_idx])
        denoised = ops.combine3(sample.contiguous(), c1, model_output.contiguous(), c2)
        if t_idx == self.num_inference_steps - 1:
            return denoised, denoised
        a_prev = self.alpha_at(self._ts_host[t_idx + 1])
        if noise is None:
            noise = torch.randn_like(sample)
        prev = ops.combine3(denoised, a_prev ** 0.5, noise.contiguous(), (1.0 - a_prev) ** 0.5)
        return prev, denoised
