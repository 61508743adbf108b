"""Shared test drivers (used by tests/ and by tests/golden/make_golden.py)."""
import torch


def host_quantile_without_contraction(real_quantile):
    """A stand-in for torch.quantile in the GPU parity tests.  The checks there read "the kernel's threshold IS torch.quantile of the
    kernel's own map, bit for bit", computed on the host.  torch.quantile's last step is a lerp whose multiply-add torch contracts into
    an FMA on some hosts (AVX2 / AVX-512 dispatch of its CPU kernels) and not on others, one unit in the last place apart (oracle.lerp_torch);
    the kernels' default is the uncontracted form.  For a 2-D fp32 CPU input reduced over dim 1 with a scalar q this returns the oracle's
    exact restatement with the uncontracted lerp, after checking that the host's own torch.quantile is within that last bit of it; every
    other call goes to torch unchanged (CUDA inputs in particular).  On a host that does not contract this is the identity."""
    from oracle import du_oracle as O

    def ulp_line(x):
        i = x.contiguous().view(torch.int32).long()
        return torch.where(i < 0, -(i & 0x7FFFFFFF), i)

    def quantile(input, q, dim=None, keepdim=False, **kw):
        ref = real_quantile(input, q, dim=dim, keepdim=keepdim, **kw)
        if (kw or not isinstance(q, (int, float)) or not torch.is_tensor(input) or input.device.type != "cpu" or input.dtype != torch.float32
                or input.dim() != 2 or dim not in (1, -1) or input.shape[0] == 0):
            return ref
        thr = O.quantile_linear_rows(input, float(q), lerp_fma=False)[0]
        flat = ref.reshape(-1)
        nan = torch.isnan(thr)
        assert torch.equal(nan, torch.isnan(flat)), "NaN rows differ between torch.quantile and its restatement"
        if not bool(nan.all()):
            assert int((ulp_line(thr[~nan]) - ulp_line(flat[~nan])).abs().max()) <= 1, "torch.quantile is not within the lerp's last bit"
        return thr.reshape(ref.shape)

    return quantile


def l4_sampling_loop(scheduler, model, x_T, y, eta=0.0, slice_channels=None):
    """The reference's L4 timestep loop around scheduler.step(), restated:
    diffusion_uncertainty/generate_samples.py:175-201 (class-conditioned, from tensor).
    Returns dict(final, uncertainty [B,T_uc,...] or None, score [B,T_uc,...] or None, prevs list)."""
    x = x_T
    C = slice_channels if slice_channels is not None else x_T.shape[1]
    scheduler.prompt_embeds = y
    uncs, scores, prevs = [], [], []
    with torch.no_grad():
        for t in scheduler.timesteps:
            t = int(t.item())
            t_tensor = torch.full((x.shape[0],), t, device=x.device, dtype=torch.long)
            x = scheduler.scale_model_input(x, t)
            eps = model(x, t_tensor, y=y)[:, :C]
            out = scheduler.step(eps, t, x, eta=eta) if eta != 0.0 else scheduler.step(eps, t, x)   # (the DPM-2 step has no eta)
            if scheduler.timestep_after_step >= t >= scheduler.timestep_end_step:
                uncs.append(out.uncertainty.detach().cpu())
                pe = out.pred_epsilon
                scores.append(pe.detach().cpu())
            x = out.prev_sample
            prevs.append(x.detach().cpu())
    return {
        "final": x.detach().cpu(),
        "uncertainty": torch.stack(uncs, dim=1) if uncs else None,
        "score": torch.stack(scores, dim=1) if scores else None,
        "prevs": prevs,
    }


def l4_unconditioned_loop(scheduler, model, X_T, batch_size):
    """The reference's unconditioned loop restated (diffusion_uncertainty/generate_samples.py:366-463): batches of X_T,
    `model(x, t).sample[:, :3]`, scheduler.step, per-step maps / scores stacked on dim 1, uint8 epilogue."""
    imgs, uncs, scores = [], [], []
    with torch.no_grad():
        for a in range(0, X_T.shape[0], batch_size):
            x = X_T[a:a + batch_size]
            u_b, s_b = [], []
            for t in scheduler.timesteps:
                t = int(t.item())
                t_tensor = torch.full((x.shape[0],), t, device=x.device, dtype=torch.long)
                eps = model(x, t_tensor).sample[:, :3]
                out = scheduler.step(eps, t, x)
                if scheduler.timestep_after_step >= t >= scheduler.timestep_end_step:
                    u_b.append(out.uncertainty.cpu())
                    s_b.append(out.pred_epsilon.cpu())
                x = out.prev_sample
            uncs.append(torch.stack(u_b, dim=1))
            scores.append(torch.stack(s_b, dim=1))
            imgs.append(((x / 2 + 0.5).clamp(0, 1) * 255.0).round().to(torch.uint8))
    return {"gen_images": torch.cat(imgs, 0).cpu(), "uncertainty": torch.cat(uncs, 0), "score": torch.cat(scores, 0)}


def l4_uvit_loop(scheduler, uvit_ae, X_T, y, batch_size):
    """The reference's U-ViT latent loop restated (diffusion_uncertainty/generate_samples.py:469-571)."""
    imgs, uncs, scores = [], [], []
    with torch.no_grad():
        for a in range(0, X_T.shape[0], batch_size):
            x, yb = X_T[a:a + batch_size], y[a:a + batch_size]
            scheduler.prompt_embeds = yb
            scheduler.set_timesteps(len(scheduler.timesteps))
            u_b, s_b = [], []
            for t in scheduler.timesteps:
                t = int(t.item())
                t_tensor = torch.full((yb.shape[0],), t, device=x.device, dtype=torch.long)
                out = scheduler.step(uvit_ae(x, t_tensor, yb), t, x)
                if scheduler.timestep_after_step >= t >= scheduler.timestep_end_step:
                    u_b.append(out.uncertainty.cpu())
                    s_b.append(out.pred_epsilon.cpu())
                x = out.prev_sample
            uncs.append(torch.stack(u_b, dim=1))
            scores.append(torch.stack(s_b, dim=1))
            x = uvit_ae.decode(x)
            imgs.append(((x / 2 + 0.5).clamp(0, 1) * 255.0).round().to(torch.uint8))
    return {"gen_images": torch.cat(imgs, 0).cpu(), "uncertainty": torch.cat(uncs, 0), "score": torch.cat(scores, 0)}
