"""Shared test drivers (used by tests/ and by tests/golden/make_golden.py)."""
import torch


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
