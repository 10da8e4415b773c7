"""Oracle restatement of the closure ``conditon_function`` (clip_diffusion/sample.py:134-238), TEST-ONLY, CPU fp32.

Line-for-line the reference's control flow and autograd calls, on the oracle operators (oracle.cutouts /
oracle.losses / oracle.clip_vit).  Differences, all needed to make it a checker:
  * random decisions come from ``record_source`` (explicit RNG records incl. noise) instead of the global
    generator, so the CUDA path can be fed the identical parameters;
  * ``diffusion`` / ``model`` are arguments (any object with p_mean_variance / sqrt_one_minus_alphas_cumprod);
  * the init-image branch (sample.py:220-225): the MS-SSIM term is restated (oracle/ms_ssim.py, parity unpinned: ``pytorch_msssim`` is
    not vendored); the LPIPS term needs the un-vendored VGG network and is left out.
"""
import torch

from oracle import cutouts as OC
from oracle import losses as OL


def make_conditon_function(diffusion, model, clip_models, text_embeddings_and_weights, get_current_timestep, config, record_source,
                           aesthetic_predictors=None, range_scale=0.0, init_image_tensor=None):
    aesthetic_predictors = aesthetic_predictors or {}

    @torch.enable_grad()
    def conditon_function(x, t, y=None):
        x = x.detach().requires_grad_()  # sample.py:143
        batch_size = x.shape[0]
        current_timestep = get_current_timestep()
        ts = torch.ones([batch_size], dtype=torch.long) * current_timestep  # :146-148
        p_mean_var = diffusion.p_mean_variance(model, x, ts, clip_denoised=False, model_kwargs={"y": y})  # :149-151
        factor = float(diffusion.sqrt_one_minus_alphas_cumprod[current_timestep])  # :152
        denoised_prediction = p_mean_var["pred_xstart"] * factor + x * (1 - factor)  # :154
        grad_tensor = torch.zeros_like(denoised_prediction)  # :155
        current_diffusion_step = 1000 - (int(t.item()) + 1)  # :157-159
        n_over = config.num_overview_cuts_schedule[current_diffusion_step]
        n_inner = config.num_inner_cuts_schedule[current_diffusion_step]
        H, W = denoised_prediction.shape[2:4]
        for name, clip_model in clip_models.items():  # :161
            for b in range(config.num_cutout_batches):  # :162
                rec = record_source(name, b, H, W, clip_model.visual.input_resolution, n_over, n_inner,
                                    config.inner_cut_size_power_schedule[current_diffusion_step],
                                    config.cut_gray_portion_schedule[current_diffusion_step])
                cutout_images = OC.make_cutouts(denoised_prediction, rec)  # :165-172
                image_embeddings = clip_model.encode_image(OC.clip_normalize(cutout_images)).float()  # :173
                aesthetic_score = None
                if config.aesthetic_scale > 0 and name in aesthetic_predictors:  # :175-176
                    aesthetic_score = OL.aesthetic_loss(aesthetic_predictors[name], image_embeddings)
                distances = OL.square_spherical_distance_loss(  # :179-182
                    image_embeddings.unsqueeze(1), text_embeddings_and_weights[name]["embeddings"].unsqueeze(0))
                distances = distances.view([n_over + n_inner, batch_size, -1])  # :184-191
                distance_loss = distances.mul(text_embeddings_and_weights[name]["weights"]).sum(dim=2).mean(dim=0)  # :194-198
                objective = distance_loss.sum() * config.clip_guidance_scale
                if aesthetic_score is not None:
                    objective = objective - aesthetic_score * config.aesthetic_scale  # :202-203
                grad_tensor += torch.autograd.grad(objective, denoised_prediction)[0] / config.num_cutout_batches  # :199-214
        loss_sum = OL.total_variational_loss(denoised_prediction).sum() * config.denoise_scale  # :217-218
        if range_scale:
            loss_sum = loss_sum + OL.rgb_range_loss(denoised_prediction).sum() * range_scale
        if init_image_tensor is not None:  # :220-225 (dissimlarity_loss.sum() * Config.MS_SSIM_scale)
            from oracle.ms_ssim import structural_dissimilarity_loss

            loss_sum = loss_sum + structural_dissimilarity_loss(denoised_prediction, init_image_tensor).sum() * config.MS_SSIM_scale
        grad_tensor += torch.autograd.grad(loss_sum, denoised_prediction)[0]  # :226
        conditon_function.last_grad_tensor = grad_tensor.detach().clone()
        if not torch.isnan(grad_tensor).any():  # :228
            grad = -torch.autograd.grad(denoised_prediction, x, grad_tensor)[0]
        else:
            return torch.zeros_like(x)
        magnitude = grad.square().mean().sqrt()  # :236
        return grad * magnitude.clamp(min=-config.grad_threshold, max=config.grad_threshold) / magnitude  # :238

    return conditon_function
