"""The evaluation drivers of image_sample.py (`evaluate_constraint` :608-710, `evaluate_unconstraint` :522-569,
`analyze_log` :584-606) on the CUDA stack: the result dictionary has the reference's keys, and its numbers equal the
oracle's restatement of the reference formulas (MSE / PSNR / L1 :674-680, ssim_fn :571-582) applied to the very images
the loop produced."""
import os
from functools import partial

import pytest
import torch

from oracle import metrics as OM
from test_gpu_constrained import _setup

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")
KW = dict(style="pred", norm_eps=True, refine_prior_sigma=True, sigma_pred_threshold=960)


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "loops3_constrained.pt"), weights_only=True)


def _loader(n_batches, B, R):
    g = torch.Generator().manual_seed(77)
    smooth = torch.nn.functional.avg_pool2d(torch.rand(n_batches * B, 3, R + 4, R + 4, generator=g), 5, 1)
    return [(smooth[i * B:(i + 1) * B].clone(), None) for i in range(n_batches)]


def test_evaluate_constraint_matches_the_reference_formulas(golden):
    from nlc_b200 import image_sample as IS
    exp, _, con = _setup("fp32", golden, "sr_averagepooling|4")[:3]
    R, B = exp.data_shape[-1], 2
    loader = _loader(3, B, R)
    torch.manual_seed(0)
    log, _ = IS.evaluate_constraint(exp, loader, con, None, **KW)
    for key in ("mse", "psner", "ssim", "const_f_loss", "const_b_loss", "const_orig_loss", "fid", "full_log", "full_results"):
        assert key in log
    assert log["fid"] is None and len(log["full_log"]["psnr"]) == 3 * B
    # the same loop by hand, scored with the oracle's formulas on the CPU
    torch.manual_seed(0)
    gen = exp.new_gen()
    mse, psnr, ssim, l1 = [], [], [], []
    for x_orig, _ in loader:
        y = con.transform((2 * x_orig - 1).to(dev))
        sample, _ = exp.denoise_loop(shape=(B, 3, R, R), gen=gen, constrain_fn=partial(con.constraint_fn, y=y, lambda_t=con.lr),
                                     constrain_loss=partial(con.loss, y=y), return_log=False, chunk_size=1, **KW)
        ref = OM.restoration_metrics(sample, x_orig)
        mse += ref["mse"].tolist()
        psnr += ref["psnr"].tolist()
        l1 += ref["const_orig"].tolist()
        ssim += OM.ssim3d(ref["image"], x_orig).tolist()
    full = log["full_log"]
    assert torch.allclose(torch.tensor(full["mse"]), torch.tensor(mse), rtol=1e-5)
    assert torch.allclose(torch.tensor(full["psnr"]), torch.tensor(psnr), rtol=1e-5)
    assert torch.allclose(torch.tensor(full["const_orig_loss"]), torch.tensor(l1), rtol=1e-5)
    assert (torch.tensor(full["ssim"]) - torch.tensor(ssim)).abs().max() < 1e-4
    assert abs(log["psner"] - sum(psnr) / len(psnr)) < 1e-4 and abs(log["ssim"] - sum(ssim) / len(ssim)) < 1e-4
    # batches are dealt round-robin to the ranks of a sharded run
    torch.manual_seed(0)
    log0, _ = IS.evaluate_constraint(exp, loader, con, None, rank=0, world=2, **KW)
    log1, _ = IS.evaluate_constraint(exp, loader, con, None, rank=1, world=2, **KW)
    assert len(log0["full_log"]["mse"]) == 2 * B and len(log1["full_log"]["mse"]) == B


def test_analyze_log_and_png_output(golden, tmp_path):
    from nlc_b200 import image_sample as IS
    exp, _, con = _setup("fp32", golden, "colorization|1")[:3]
    R, B = exp.data_shape[-1], 2
    loader = _loader(1, B, R)
    out = tmp_path / "images"
    out.mkdir()
    log, ret = IS.evaluate_constraint(exp, loader, con, str(out), return_log=True, **KW)
    assert sorted(os.listdir(out)) == ["00-00000-000.png", "00-00000-001.png"]
    res = log["full_results"][0]
    n_steps = len(ret[1])
    for key in ("zt", "z0_prec", "z0_postc"):
        assert len(res[key]["psnr"]) == n_steps and len(res[key]["ssim"]) == n_steps and len(res[key]["const"]) == n_steps
    # the projected x0 satisfies the measurement better than the raw one at every step
    assert all(a <= b + 1e-6 for a, b in zip(res["z0_postc"]["const"], res["z0_prec"]["const"]))
    # a second call finds the files and skips the batch (the reference's resume logic)
    log2, _ = IS.evaluate_constraint(exp, loader, con, str(out), **KW)
    assert log2["full_log"]["mse"] == []


def test_evaluate_unconstraint(golden):
    from nlc_b200 import image_sample as IS
    exp = _setup("fp32", golden, "colorization|1")[0]
    log, lists = IS.evaluate_unconstraint(exp, 5, None, **KW)
    s = log["samples"]
    assert s.shape == (6, 3) + tuple(exp.data_shape[-2:]) and s.min() >= 0 and s.max() <= 1 and log["fid"] is None
    logp, _ = IS.evaluate_unconstraint(exp, 2, None, sampling="project", sigma_estimate_rate=(0.5, 0.2, 0.2, 0.1), **KW)
    assert logp["samples"].shape[0] == 2 and torch.isfinite(logp["samples"]).all()


def test_evaluate_unconstraint_reports_the_device_fid(golden):
    """With fid.fid_helper the driver accumulates the InceptionV3 statistics of its own samples on the device; the reported
    FID is the Frechet distance between the target and those statistics (a 64-feature head keeps the host sqrtm small)."""
    import numpy as np
    from nlc_b200 import fid, image_sample as IS
    from oracle import fid as OF, weights
    exp = _setup("fp32", golden, "colorization|1")[0]
    sd = weights.fid_inception_state_dict(seed=7)
    net = fid.InceptionV3(precision="fp32", device=dev).load_state_dict(sd)
    seen = []
    orig = net.features_of_samples
    net.features_of_samples = lambda x: seen.append(orig(x)) or seen[-1]
    target = (np.zeros(2048), np.eye(2048))
    fid.fid_helper(exp, target, net)
    exp.fid_of = lambda st: ("stats", st.count) + st.finalize()  # (skip the 2048 x 2048 host sqrtm: return the statistics)
    log, _ = IS.evaluate_unconstraint(exp, 5, None, **KW)
    tag, count, mu, sigma = log["fid"]
    feats = torch.cat(seen).cpu().numpy()
    assert tag == "stats" and count == log["samples"].shape[0] == feats.shape[0]
    m, s = OF.statistics(feats)
    assert np.abs(mu - m).max() < 1e-9 and np.abs(sigma - s).max() < 1e-9
    # the features are those of the images the driver returned, through the 8-bit round trip
    with torch.no_grad():
        want = OF.inception_features(sd, OF.png_round_trip(log["samples"].cpu()))
    assert ((torch.from_numpy(feats) - want).abs().max() / want.abs().max()).item() < 1e-3
