"""Shared helpers of the GPU parity tests: per-sample errors and time-bucket flips.

The NLC step evaluates the UNet at t_hat = searchsorted(sigma_table, sigma_hat) (src/schedulers.py:185-190).  The table
has ~1 % bucket spacing, so in the reduced-precision modes (sigma_hat error of a few 1e-4) a sample whose sigma_hat lies
close to a bucket edge gets the neighbouring integer t_hat; with random-init weights that moves its eps by ~1e-2.  The
teacher-forced tests therefore judge every sample on its own: the stated tolerance when its t_hat equals the
reference's, a bounded one (FLIP_TOL) when it is off by exactly one bucket; anything else fails."""
import torch

FLIP_TOL = 1e-1


def per_sample_l2rel(a, b):
    a, b = a.double().flatten(1), b.double().flatten(1)
    return torch.linalg.vector_norm(a - b, dim=1) / torch.linalg.vector_norm(b, dim=1).clamp_min(1e-30)


def bucket_distance(sch, sigma_ours, sigma_ref):
    """|t_hat(ours) - t_hat(reference)| per sample (0 everywhere in continuous-t mode)."""
    if getattr(sch, "continuous_t", False):
        return torch.zeros(sigma_ref.numel(), dtype=torch.long)
    dev = sch.sigmas.device
    to = sch.sigma_to_t(sigma_ours.reshape(-1).to(dev).float())
    tr = sch.sigma_to_t(sigma_ref.reshape(-1).to(dev).float())
    return (to - tr).abs().cpu()


def assert_step_close(name, ours, ref, tol, dist, context):
    """Every sample within `tol`, or within FLIP_TOL if its time bucket is the reference's neighbour."""
    err = per_sample_l2rel(ours, ref)
    lim = torch.where(dist == 0, torch.full_like(err, tol), torch.full_like(err, max(tol, FLIP_TOL)))
    assert (dist <= 1).all(), (context, name, "t_hat more than one bucket off", dist.tolist())
    assert (err < lim).all(), (context, name, err.tolist(), lim.tolist())


# ------------------------------------------------------------------------------------------------------------------------
# The reference's own GPU arithmetic as a control for free-running trajectories.
def oracle_c2_loop_on_gpu(g, z, noises, tf32=True, autocast=None, forced=False, dev="cuda:0"):
    """Config c2's 100-step NLC loop (batch 4, tests/golden/loop_c2_100.pt) through the ORACLE - pinned torch.equal to the
    reference on the CPU - on the GPU with PyTorch eager / cuDNN: `tf32=True` is the reference's default GPU path (TF32
    convolutions), `tf32=False` strict fp32 cuDNN, `autocast` a torch.autocast dtype.  `forced`: both time lookups
    t = searchsorted(sigma) of every step are taken from the recorded CPU run.  Returns (final image on the CPU, number of
    the 4 samples that crossed a time-bucket edge).  Test infrastructure: what "matches the reference" can mean for a
    reduced-precision run is bounded by how well the reference's own GPU run matches its CPU run."""
    from oracle import ddim_net, sampler as S, weights
    dev = torch.device(dev)
    cfg = weights.CONFIGS["c2"]
    sd = {k: v.to(dev) for k, v in weights.ddim_unet_state_dict(**cfg["unet"], seed=3).items()}
    ssd = {k: v.to(dev) for k, v in weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4).items()}
    tab = S.Tables()
    ts, sig, mvc = tab.ddim_schedule(100.0, None, 100)
    assert torch.equal(ts, g["timesteps"]) and torch.equal(sig, g["sigmas"])
    for name in ("betas", "alphas_cumprod", "sigmas", "posterior_variance"):
        setattr(tab, name, getattr(tab, name).to(dev))
    mvc = mvc.to(dev) if torch.is_tensor(mvc) else mvc
    d = 3 * 64 * 64
    xT = (z / (1 / (g["sigmas"][0] ** 2 + 1)).sqrt()).to(dev)
    noises_d = [n.to(dev) for n in noises]

    class ForcedTables:
        def __init__(self, inner):
            self.inner, self.calls = inner, 0

        def __getattr__(self, k):
            return getattr(self.inner, k)

        def t_of_sigma(self, sigma):  # two lookups per step: the refined sigma (encode pass), the corrected one (forward)
            i, which = divmod(self.calls, 2)
            self.calls += 1
            return (g["t_first"][i] if which == 0 else g["t_hat"][i]).to(dev)

    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32, False
    try:
        log = []
        # (`with torch.device`: the oracle creates a few tensors without a device argument; a context manager, unlike
        # torch.set_default_device, leaves no device mode behind for the tests that follow)
        with torch.device(dev), torch.no_grad(), torch.autocast("cuda", dtype=autocast or torch.float16,
                                                                enabled=autocast is not None):
            out = S.denoise_loop(ForcedTables(tab) if forced else tab, ts.tolist(), sig.to(dev), mvc,
                                 lambda z_, t: ddim_net.unet_forward(sd, z_, t), lambda z_, t: ddim_net.unet_encode(sd, z_, t),
                                 lambda f: ddim_net.sigma_forward(ssd, f), xT, kind="ddim_simple_orig", eta=0.85, style="pred",
                                 norm_eps=True, refine=True, norm_min=-2.0 / d ** 0.5, norm_max=110.0 / d ** 0.5,
                                 noises=noises_d, sigma_pred_threshold=960, log=log)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    sig_log = torch.stack([s["sigma_t"].reshape(-1).expand(4) for s in log]).float().cpu()
    flips = int((torch.searchsorted(g["table"], sig_log.contiguous()) != g["t_hat"]).any(dim=0).sum())
    return out.float().cpu(), flips
