"""Host-side mirror of the evaluation drivers of the reference's image_sample.py: `evaluate_constraint` (:608-710),
`evaluate_unconstraint` (:522-569), `ssim_fn` (:571-582), `analyze_log` (:584-606) and the module-level
`projection_loop` (:431-519), with the same arguments and the same result dictionary keys.

What changes underneath: the sampling loops are the nlc_b200 ones (`ExperimentDiffusion.denoise_loop / projection_loop`),
and the per-batch metrics (MSE, PSNR, SSIM, constraint residuals) are computed from the device tensors by libnlc_b200
kernels (`metrics.restoration_metrics`) instead of the reference's CPU / numpy / per-image cuDNN round trip.  Batches are
dealt round-robin to the ranks of a sharded run (`rank`, `world`), metric means are all-reduced at the end; with one
process the numbers are the reference's.  PNG output is kept (optional: `images_dir=None` skips it and the reference's
resume-by-existing-files logic); FID statistics are accumulated on the device (nlc_b200.fid) when the experiment has been
given a target with `fid.fid_helper`, else taken from a caller-supplied `experiment.fid_fn`, else reported as None."""
import math
import os
from functools import partial
from time import time

import numpy as np
import torch

from . import metrics as M


def ssim_fn(sample, orig):
    """image_sample.py:571-582: per-image SSIM (uint8 rounding, basicsr 3-D window) as a Python list."""
    return M.ssim_fn(sample, orig).cpu().tolist()


def projection_loop(self, *args, **kwargs):
    """image_sample.py:431-519 takes the experiment as `self`; the implementation lives on the experiment class."""
    return self.projection_loop(*args, **kwargs)


def _save_batch(sample01, images_dir, rank, i):
    if images_dir is None:
        return
    from torchvision.utils import save_image
    for j, img in enumerate(sample01):
        save_image(img, os.path.join(images_dir, f"{rank:02}-{i:05}-{j:03}.png"))


def _already_done(images_dir, rank, i, batch_size):
    if images_dir is None:
        return False
    return all(os.path.exists(os.path.join(images_dir, f"{rank:02}-{i:05}-{j:03}.png")) for j in range(batch_size))


class _Fid:
    """FID of a driver run.  With `fid.fid_helper(experiment, fid_target, inception)` the InceptionV3 statistics are
    accumulated on the device batch by batch from the sampler outputs (through the 8-bit image round trip the reference's
    PNG files go through) and all-reduced over the ranks at the end; an `experiment.fid_fn` supplied by the caller instead
    is called on `images_dir` as the reference does (src/experiments.py:220-226); without either the FID is None."""

    def __init__(self, experiment, images_dir):
        self.experiment, self.images_dir = experiment, images_dir
        make = getattr(experiment, "fid_stats", None)
        self.stats = make() if make is not None else None

    def update(self, sample_pm1):
        if self.stats is not None:
            net = self.experiment.fid_inception
            self.stats.update(net.features_of_samples(sample_pm1.to(net.device, torch.float32)))

    def value(self):
        if self.stats is not None:
            return self.experiment.fid_of(self.stats)
        fn = getattr(self.experiment, "fid_fn", None)
        return fn(self.images_dir) if (fn is not None and self.images_dir is not None) else None


class BatchStreams:
    """Noise streams of a run whose batches are dealt round-robin to `world` ranks.

    The un-sharded reference draws everything from two sequential streams: x_T of batch i is the i-th draw of the CPU
    generator `experiment.new_gen()` (src/experiments.py:260-271) and every step's z comes from the device's default
    generator (src/schedulers.py:438...).  A rank that simply started both streams at their beginning would hand its
    first batch the noise of batch 0 - every rank the same images.  Here each rank REPLAYS the global streams and
    discards the draws that belong to batches it does not own (the policy of parallel.sharded_noise), so - with every
    rank seeded alike, as torch.manual_seed / torch.cuda.manual_seed_all(seed) do - the union of the shards is the
    un-sharded run, sample for sample.  (A NaN early exit, which shortens one batch's draws, is not replayed.)"""

    def __init__(self, experiment, shape, rank=0, world=1, device_draws_per_batch=0, host_draws_per_batch=1):
        self.shape, self.rank, self.world = tuple(shape), rank, world
        self.device = experiment.device
        self.gen = experiment.new_gen()
        self.n_dev, self.n_host = device_draws_per_batch, host_draws_per_batch
        self.next_batch = 0

    def advance_to(self, i):
        """Discard the draws of batches [next_batch, i); returns the CPU generator positioned at batch i."""
        for _ in range(self.next_batch, i):
            for _ in range(self.n_host):
                torch.randn(self.shape, generator=self.gen)
            for _ in range(self.n_dev):
                torch.randn(self.shape, device=self.device)
        self.next_batch = i + 1
        return self.gen


def device_draws_per_batch(experiment, sampling="denoise", max_T=None, new_eta=None):
    """How many torch.randn_like(x0) calls one batch makes on the device generator (Scheduler.pred_xprev draws iff the
    scheduler is a DDPM kind or eta > 0; `new_eta` replaces eta for the last step, src/experiments.py:347-348)."""
    sch = experiment.scheduler
    steps = len(sch.timesteps_host) - 1 if hasattr(sch, "timesteps_host") else len(sch.timesteps) - 1
    if sampling == "project" and max_T is not None:
        steps = max_T
    always = sch.kind in ("ddpm", "ddpm_orig")
    n = steps if (always or float(sch.eta) > 0) else 0
    if new_eta is not None and not always:
        last = 1 if float(new_eta) > 0 else 0
        first = steps - 1 if float(sch.eta) > 0 else 0
        n = first + last
    return n


def evaluate_unconstraint(experiment, n_samples, images_dir, norm_init_noise=False, style="base", sampling="denoise",
                          norm_eps=False, refine_prior_sigma=False, sigma_estimate_rate=(1, 0, 0), max_T=None,
                          sigma_pred_threshold=1000, new_eta=None, recal_sigma_prev=False, return_log=False,
                          res_pkl_path="", rank=0, world=1):
    """image_sample.py:522-569.  Returns (log_dict, return_lists); `log_dict['samples']` additionally holds this rank's
    images in [0, 1] on the device (the reference only writes them to disk)."""
    batch_size = experiment.batch_size
    n_batches = math.ceil(n_samples / batch_size)
    shape = (batch_size,) + tuple(experiment.data_shape)
    streams = BatchStreams(experiment, shape, rank, world,
                           device_draws_per_batch(experiment, sampling, max_T, new_eta))
    return_lists, kept = [], []
    fid = _Fid(experiment, images_dir)
    for i in range(rank, n_batches, world):
        if _already_done(images_dir, rank, i, batch_size):
            continue
        gen = streams.advance_to(i)
        if sampling == "project":
            sample, return_list = experiment.projection_loop(
                shape=shape, gen=gen, norm_init_noise=norm_init_noise, style=style, constrain_fn=None, norm_eps=norm_eps,
                refine_prior_sigma=refine_prior_sigma, xT=None, return_log=return_log, chunk_size=1,
                sigma_estimate_rate=sigma_estimate_rate, constrain_loss=None, max_T=max_T, stop_condition=0.0,
                sigma_pred_threshold=sigma_pred_threshold, new_eta=new_eta, recal_sigma_prev=recal_sigma_prev,
                to_cpu=False)
        else:
            sample, return_list = experiment.denoise_loop(
                shape=shape, gen=gen, norm_init_noise=norm_init_noise, style=style, constrain_fn=None, norm_eps=norm_eps,
                refine_prior_sigma=refine_prior_sigma, return_log=return_log, chunk_size=1,
                sigma_pred_threshold=sigma_pred_threshold, new_eta=new_eta, to_cpu=False)
        return_lists.append(return_list)
        if return_log and res_pkl_path:
            import joblib
            joblib.dump(return_lists, res_pkl_path)
        fid.update(sample)
        sample = sample.add(1).div(2).clamp(0, 1)
        _save_batch(sample, images_dir, rank, i)
        kept.append(sample)
    log_dict = {"fid": fid.value(), "samples": torch.cat(kept) if kept else None}
    return log_dict, return_lists


def analyze_log(return_list, x_orig, y, constrain_loss):
    """image_sample.py:584-606: PSNR / SSIM / constraint loss of x_t, x0 before and after the projection, per step."""
    z_list, eps_list, x0_prec_list, x0_postc_list = return_list[:4]
    keys = ["zt", "z0_prec", "z0_postc"]
    res = {k: {"psnr": [], "ssim": [], "const": []} for k in keys}
    dev = y.device
    x_orig = x_orig.to(dev)
    for kk in range(len(eps_list)):
        for key, sample in zip(keys, (z_list[kk], x0_prec_list[kk], x0_postc_list[kk])):
            m = M.restoration_metrics(sample.to(dev), x_orig, ssim=True, return_image=True)
            mse = m["mse"].mean()  # equal image sizes: the mean of per-image means is the batch mean (:595)
            const, _ = constrain_loss(2 * m["image"] - 1.0, y)
            res[key]["psnr"].append((10 * torch.log10(1 / mse)).item())
            res[key]["ssim"].append(float(m["ssim"].mean()))
            res[key]["const"].append(float(torch.mean(const)))
    return res


@torch.no_grad()
def evaluate_constraint(experiment, data_loader, Constraint, images_dir, n_samples=-1, transform_dir=None,
                        norm_init_noise=False, style="base", sampling="denoise", norm_eps=False,
                        refine_prior_sigma=False, prior_xt=False, sigma_estimate_rate=(1, 0, 0), return_log=False,
                        max_T=None, sigma_pred_threshold=1000, new_eta=None, recal_sigma_prev=False, rank=0, world=1):
    """image_sample.py:608-710: restore every batch of `data_loader` (ground truth in [0, 1]) under `Constraint`, score it.
    Returns (log_dict, return_list) with the reference's keys (`mse`, `psner` [sic], `ssim`, `const_f_loss`,
    `const_b_loss`, `const_orig_loss`, `fid`, `full_log`, `full_results`); the means are over all ranks' samples."""
    device = experiment.device
    streams = None
    lists = {k: [] for k in ("mse", "psnr", "ssim", "const_f", "const_b", "const_orig")}
    full_results, return_list = [], None
    fid = _Fid(experiment, images_dir)
    for i, (x_orig, _classes) in enumerate(data_loader):
        if i % world != rank:
            continue
        batch_size = x_orig.shape[0]
        if streams is None:
            # x_T is one CPU draw per batch, or (prior_xt) one more device draw; then the per-step draws
            n_dev = device_draws_per_batch(experiment, sampling, max_T, new_eta) + (1 if prior_xt else 0)
            streams = BatchStreams(experiment, (batch_size,) + tuple(experiment.data_shape), rank, world, n_dev,
                                   host_draws_per_batch=0 if prior_xt else 1)
        gen = streams.advance_to(i)
        x_orig = x_orig.to(device)
        batch_x = 2 * x_orig - 1.0
        if _already_done(images_dir, rank, i, batch_size):
            continue
        y = Constraint.transform(batch_x)
        Apy = None
        if transform_dir is not None or prior_xt:
            Apy = Constraint.inv_transform(y)
        if transform_dir is not None:
            from torchvision.utils import save_image
            sample_apy = Apy.add(1).div(2).clamp(0, 1)
            for j in range(len(Apy)):
                save_image(sample_apy[j], os.path.join(transform_dir, f"Apy_{rank:02}-{i:05}-{j:03}.png"))
                save_image(x_orig[j], os.path.join(transform_dir, f"orig_{rank:02}-{i:05}-{j:03}.png"))
        shape = (batch_size,) + tuple(experiment.data_shape)
        constraint_fn = partial(Constraint.constraint_fn, y=y, lambda_t=Constraint.lr)
        constrain_loss = partial(Constraint.loss, y=y)
        xT = Apy + experiment.scheduler.sampling_sigmas[0] * torch.randn_like(Apy) if prior_xt else None
        t1 = time()
        if sampling == "project":
            sample, return_list = experiment.projection_loop(
                shape=shape, gen=gen, norm_init_noise=norm_init_noise, style=style, constrain_fn=constraint_fn,
                norm_eps=norm_eps, refine_prior_sigma=refine_prior_sigma, xT=xT, return_log=return_log, chunk_size=1,
                sigma_estimate_rate=sigma_estimate_rate, constrain_loss=constrain_loss, max_T=max_T, stop_condition=0.0,
                sigma_pred_threshold=sigma_pred_threshold, new_eta=new_eta, recal_sigma_prev=recal_sigma_prev,
                to_cpu=False)
        else:
            sample, return_list = experiment.denoise_loop(
                shape=shape, gen=gen, norm_init_noise=norm_init_noise, style=style, constrain_fn=constraint_fn,
                norm_eps=norm_eps, refine_prior_sigma=refine_prior_sigma, xT=xT, return_log=return_log, chunk_size=1,
                constrain_loss=constrain_loss, sigma_pred_threshold=sigma_pred_threshold, new_eta=new_eta, to_cpu=False)
        elapsed = time() - t1
        m = M.restoration_metrics(sample.to(device), x_orig, constraint=Constraint, y=y, return_image=True, ssim=True)
        fid.update(sample)
        _save_batch(m["image"], images_dir, rank, i)
        for k in lists:
            lists[k] += m[k].cpu().tolist()
        print(f"done batches:{i}/{len(data_loader)}, time:{elapsed:.2f}  psnr:{np.mean(lists['psnr'])}, "
              f"ssim:{np.mean(lists['ssim'])}, cost:{np.mean(lists['const_f'])}")
        if return_log:
            full_results.append(analyze_log(return_list, x_orig, y, Constraint.loss))
        if n_samples > 0 and (i + 1) * batch_size > n_samples:
            break
    # global means over the samples of all ranks (one all-reduce of the sums and the count)
    means = M.reduce_means({k: torch.tensor(v, dtype=torch.float64, device=device) for k, v in lists.items()},
                           keys=tuple(lists))
    log_dict = {"mse": means["mse"], "psner": means["psnr"], "ssim": means["ssim"], "const_f_loss": means["const_f"],
                "const_b_loss": means["const_b"], "const_orig_loss": means["const_orig"],
                "fid": fid.value(),
                "full_log": {"psnr": lists["psnr"], "mse": lists["mse"], "ssim": lists["ssim"],
                             "const_forward": lists["const_f"], "const_backward": lists["const_b"],
                             "const_orig_loss": lists["const_orig"]},
                "full_results": full_results}
    return log_dict, return_list
