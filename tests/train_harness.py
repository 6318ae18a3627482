"""Caller-side mirror of ONE GS-SR 2DGS training iteration, used to exercise the drop-in rasterizer exactly
the way GS-SR drives it and to measure train iters/s (BASELINE metric, second half).

It restates, with plain torch ops, what the reference does around the rasterizer call (the reference's own
Python cannot travel to the GPU box):
  Trainer.train step                     gssr/engine/trainer.py:88-128 (loss.backward, loss.item, optimizer.step, zero_grad)
  activations of the raw parameters      gssr/gaussian/vanilla_gaussian.py (exp / sigmoid / normalize, get_features cat)
  TwoDGSScene.render                     gssr/scene/twodgs_scene.py:37-127 (means2D protocol :39-43, allmap post-processing :88-117)
  depth_to_normal / depths_to_points     gssr/utils/point_utils.py:9-37
  VanillaScene L1 + SSIM                 gssr/scene/vanilla_scene.py:29-69 (lambda_dssim 0.2)
  TwoDGSScene normal / dist losses       gssr/scene/twodgs_scene.py:25-35 (lambda_normal 0.05, lambda_dist as configured)
  densification statistics               gssr/gaussian/vanilla_gaussian.py:428-430 (viewspace grad norm, denom, max radii)
The rasterizer is pluggable: the drop-in ``diff_surfel_rasterization`` (product) or the UNMODIFIED reference
kernels (oracle/_ref/libref_surfel.so) behind an autograd.Function -- everything else is shared, so the iters/s
ratio isolates the rasterizer.  Test infrastructure, not product code.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

import synth


class _RefSurfelFn(torch.autograd.Function):
    """Reference CUDA kernels as an autograd op with the same signature/returns as the reference
    _RasterizeGaussians (S/diff_surfel_rasterization/__init__.py:44-156)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors, opacities, scales, rotations, ref, rs):
        color, radii, others, _ = ref.forward(rs.bg, rs.viewmatrix, rs.projmatrix, rs.campos, rs.image_width,
                                              rs.image_height, rs.tanfovx, rs.tanfovy, means3D.contiguous(),
                                              opacities.contiguous(), scales.contiguous(), rotations.contiguous(),
                                              shs=None if sh is None else sh.contiguous(),
                                              colors=None if colors is None else colors.contiguous(),
                                              sh_degree=rs.sh_degree)
        ctx.ref, ctx.has_sh = ref, sh is not None
        ctx.mark_non_differentiable(radii)
        return color, radii, others

    @staticmethod
    def backward(ctx, g_color, g_radii, g_others):
        g = ctx.ref.backward(g_color, g_others)
        return (g["means3D"], g["means2D"], g["shs"] if ctx.has_sh else None, None if ctx.has_sh else g["colors"],
                g["opacities"], g["scales"], g["rotations"], None, None)


class MiniTwoDGSTrainer:
    def __init__(self, P=100_000, W=800, H=800, seed=0, impl="ours", lambda_dssim=0.2, lambda_normal=0.05,
                 lambda_dist=100.0, depth_ratio=0.0, device="cuda", capturable=False):
        self.impl, self.W, self.H, self.device = impl, W, H, device
        self.lambda_dssim, self.lambda_normal, self.lambda_dist, self.depth_ratio = lambda_dssim, lambda_normal, lambda_dist, depth_ratio
        sc = synth.make_scene(P, W, H, seed=seed, sh=True)
        self.sc = sc
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
        cam = sc.cam
        self.view, self.proj, self.campos, self.bg = t(cam.viewmatrix), t(cam.projmatrix), t(cam.campos), t(cam.bg)
        # raw parameters as GS-SR stores them (log-scales, logit-opacities, un-normalised quats, SH split dc / rest)
        self.xyz = torch.nn.Parameter(t(sc.means3D))
        self.features_dc = torch.nn.Parameter(t(sc.shs[:, :1, :]))
        self.features_rest = torch.nn.Parameter(t(sc.shs[:, 1:, :]))
        self.scaling = torch.nn.Parameter(torch.log(t(sc.scales)))
        self.rotation = torch.nn.Parameter(t(sc.rotations) * 1.7)
        op = t(sc.opacities).clamp(1e-4, 1 - 1e-4)
        self.opacity = torch.nn.Parameter(torch.log(op / (1 - op)))
        self.optimizer = torch.optim.Adam([
            {"params": [self.xyz], "lr": 1.6e-4, "name": "xyz"},
            {"params": [self.features_dc], "lr": 2.5e-3, "name": "f_dc"},
            {"params": [self.features_rest], "lr": 2.5e-3 / 20.0, "name": "f_rest"},
            {"params": [self.opacity], "lr": 0.05, "name": "opacity"},
            {"params": [self.scaling], "lr": 5e-3, "name": "scaling"},
            {"params": [self.rotation], "lr": 1e-3, "name": "rotation"}], lr=0.0, eps=1e-15, capturable=capturable)
        self.xyz_gradient_accum = torch.zeros((P, 1), device=device)
        self.denom = torch.zeros((P, 1), device=device)
        self.max_radii2D = torch.zeros((P,), device=device)
        # a fixed "ground truth" image (smooth pattern): the same for both arms
        yy, xx = torch.meshgrid(torch.arange(H, device=device).float() / H, torch.arange(W, device=device).float() / W, indexing="ij")
        self.gt = torch.stack([0.5 + 0.4 * torch.sin(6 * xx + 2 * yy), 0.5 + 0.4 * torch.cos(5 * yy), 0.5 + 0.4 * torch.sin(4 * (xx - yy))])
        g1 = torch.tensor([math.exp(-(x - 5) ** 2 / (2 * 1.5 ** 2)) for x in range(11)])
        g1 = (g1 / g1.sum()).unsqueeze(1)
        self.window = g1.mm(g1.t()).float()[None, None].expand(3, 1, 11, 11).contiguous().to(device)
        if impl == "ours":
            from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
            self.Settings, self.Rasterizer = GaussianRasterizationSettings, GaussianRasterizer
        else:
            from diff_surfel_rasterization import GaussianRasterizationSettings   # same NamedTuple as the reference's
            from oracle.refcuda import RefSurfel
            self.Settings, self.ref = GaussianRasterizationSettings, RefSurfel()

    # ---- twodgs_scene.py:37-127 ---------------------------------------------------------------------------
    def render(self):
        means3D = self.xyz
        opacity = torch.sigmoid(self.opacity)
        scales = torch.exp(self.scaling)
        rotations = F.normalize(self.rotation)
        shs = torch.cat((self.features_dc, self.features_rest), dim=1)
        screenspace_points = torch.zeros_like(means3D, dtype=means3D.dtype, requires_grad=True, device=self.device) + 0
        screenspace_points.retain_grad()
        rs = self.Settings(image_height=self.H, image_width=self.W, tanfovx=self.sc.cam.tanfovx, tanfovy=self.sc.cam.tanfovy,
                           bg=self.bg, scale_modifier=1.0, viewmatrix=self.view, projmatrix=self.proj, sh_degree=3,
                           campos=self.campos, prefiltered=False, debug=False)
        if self.impl == "ours":
            rendered_image, radii, allmap = self.Rasterizer(raster_settings=rs)(
                means3D=means3D, means2D=screenspace_points, shs=shs, colors_precomp=None, opacities=opacity,
                scales=scales, rotations=rotations, cov3D_precomp=None)
        else:
            rendered_image, radii, allmap = _RefSurfelFn.apply(means3D, screenspace_points, shs, None, opacity, scales,
                                                               rotations, self.ref, rs)
        if self.fused_post:
            from gsr_b200.surfel_post import surfel_postprocess
            post = surfel_postprocess(allmap, self.view, self.proj, depth_ratio=self.depth_ratio)
            return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
                    "radii": radii, **post}
        render_alpha = allmap[1:2]
        render_normal = allmap[2:5]
        render_normal = (render_normal.permute(1, 2, 0) @ (self.view[:3, :3].T)).permute(2, 0, 1)
        render_depth_median = torch.nan_to_num(allmap[5:6], 0, 0)
        render_depth_expected = torch.nan_to_num(allmap[0:1] / render_alpha, 0, 0)
        render_dist = allmap[6:7]
        surf_depth = render_depth_expected * (1 - self.depth_ratio) + self.depth_ratio * render_depth_median
        surf_normal = self.depth_to_normal(surf_depth).permute(2, 0, 1) * render_alpha.detach()
        return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
                "radii": radii, "rend_alpha": render_alpha, "rend_dist": render_dist, "surf_normal": surf_normal,
                "depth": surf_depth, "normal": render_normal}

    # ---- point_utils.py:9-37 ------------------------------------------------------------------------------
    def depth_to_normal(self, depth):
        W, H = self.W, self.H
        c2w = (self.view.T).inverse()
        ndc2pix = torch.tensor([[W / 2, 0, 0, W / 2], [0, H / 2, 0, H / 2], [0, 0, 0, 1]]).float().to(self.device).T
        projection_matrix = c2w.T @ self.proj
        intrins = (projection_matrix @ ndc2pix)[:3, :3].T
        grid_x, grid_y = torch.meshgrid(torch.arange(W, device=self.device).float(), torch.arange(H, device=self.device).float(), indexing="xy")
        points = torch.stack([grid_x, grid_y, torch.ones_like(grid_x)], dim=-1).reshape(-1, 3)
        rays_d = points @ intrins.inverse().T @ c2w[:3, :3].T
        rays_o = c2w[:3, 3]
        points = (depth.reshape(-1, 1) * rays_d + rays_o).reshape(*depth.shape[1:], 3)
        output = torch.zeros_like(points)
        dx = points[2:, 1:-1] - points[:-2, 1:-1]
        dy = points[1:-1, 2:] - points[1:-1, :-2]
        output[1:-1, 1:-1, :] = F.normalize(torch.cross(dx, dy, dim=-1), dim=-1)
        return output

    # ---- vanilla_scene.py:29-69, twodgs_scene.py:25-35 ----------------------------------------------------
    fused_ssim = False      # True: gsr_b200.ssim (fused kernel) instead of the reference's conv2d formulation
    fused_post = False      # True: gsr_b200.surfel_post (fused kernels) instead of the torch post-processing block

    def ssim(self, img1, img2):
        if self.fused_ssim:
            from gsr_b200.ssim import ssim as fused
            return fused(img1, img2)
        w, pad = self.window, 5
        mu1, mu2 = F.conv2d(img1, w, padding=pad, groups=3), F.conv2d(img2, w, padding=pad, groups=3)
        mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
        s1 = F.conv2d(img1 * img1, w, padding=pad, groups=3) - mu1_sq
        s2 = F.conv2d(img2 * img2, w, padding=pad, groups=3) - mu2_sq
        s12 = F.conv2d(img1 * img2, w, padding=pad, groups=3) - mu1_mu2
        C1, C2 = 0.01 ** 2, 0.03 ** 2
        return (((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean()

    def loss_dict(self, out):
        image = out["render"]
        d = {"L1_loss": (1.0 - self.lambda_dssim) * torch.abs(image - self.gt).mean(),
             "ssim_loss": self.lambda_dssim * (1.0 - self.ssim(image, self.gt))}
        normal_error = (1 - (out["normal"] * out["surf_normal"]).sum(dim=0))[None]
        d["normal_loss"] = self.lambda_normal * normal_error.mean()
        d["dist_loss"] = self.lambda_dist * out["rend_dist"].mean()
        return d

    # ---- trainer.py:88-128 ---------------------------------------------------------------------------------
    def step(self):
        out = self.render()
        losses = self.loss_dict(out)
        loss = sum(losses.values())
        loss.backward()
        loss_value = loss.item()                                    # the reference's per-step host sync
        with torch.no_grad():                                       # vanilla_gaussian.py:428-430 + max radii
            vis = out["visibility_filter"]
            self.max_radii2D[vis] = torch.max(self.max_radii2D[vis], out["radii"][vis].float())
            g = out["viewspace_points"].grad
            self.xyz_gradient_accum[vis] += torch.norm(g[vis, :2], dim=-1, keepdim=True)
            self.denom[vis] += 1
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=True)
        return loss_value, {k: float(v.detach()) for k, v in losses.items()}


    # ---- the same iteration without host round trips, for CUDA-graph capture ------------------------------
    def step_device(self):
        """step() with its three host interactions removed -- the per-step `loss.item()` and the two boolean-mask
        gathers/scatters of the densification statistics (data-dependent shapes) written as `torch.where` / masked adds
        (identical values) -- so that the whole iteration (render, losses, backward, statistics, Adam) can be recorded
        into ONE CUDA graph.  Only the drop-in rasterizer allows that: the reference forward blocks on a cudaMemcpy.
        Returns the loss as a device tensor; gradients are left in place (zero_grad happens before the capture)."""
        out = self.render()
        losses = self.loss_dict(out)
        loss = sum(losses.values())
        loss.backward()
        with torch.no_grad():
            vis = out["visibility_filter"]
            self.max_radii2D.copy_(torch.where(vis, torch.max(self.max_radii2D, out["radii"].float()), self.max_radii2D))
            g = out["viewspace_points"].grad
            visf = vis.unsqueeze(1).float()
            self.xyz_gradient_accum += torch.norm(g[:, :2], dim=-1, keepdim=True) * visf
            self.denom += visf
        self.optimizer.step()
        return loss.detach()

    def capture(self, warmup=3):
        """Eager warm-up on a side stream (num_rendered history, allocator, optimizer state), then one captured
        iteration.  Returns (graph, static loss tensor); every graph.replay() is one training iteration."""
        from gsr_b200.graphs import capture
        return capture(self.step_device, warmup=warmup, before_capture=lambda: self.optimizer.zero_grad(set_to_none=True))


class MiniScaffold2DGSTrainer(MiniTwoDGSTrainer):
    """Scaffold-2DGS iteration (BASELINE config 3): anchors -> scaffold_filter.visible_filter on scales[:, :3]
    (scaffold_scene.py:122-155) -> neural Gaussians from three MLPs with the opacity mask
    (scaffold_scene.py:26-120) -> surfel rasterizer with colors_precomp and the stride-3 ``scaling[:, :2]`` view
    (scaffold_2dgs_scene.py:14-19) -> 2DGS losses + scaling loss -> anchor statistics
    (scaffold_gaussian.py:488-508: opacity_accum, anchor_demon, offset_gradient_accum, offset_denom)."""

    def __init__(self, n_anchor=200_000, k=5, feat_dim=32, W=800, H=800, seed=0, impl="ours", device="cuda", **kw):
        self.impl, self.W, self.H, self.device, self.k = impl, W, H, device, k
        self.lambda_dssim, self.lambda_normal, self.lambda_dist, self.depth_ratio = 0.2, 0.05, kw.get("lambda_dist", 0.0), 0.0
        self.lambda_scaling = 0.01
        sc = synth.make_scene(n_anchor, W, H, seed=seed, scale_dims=3)
        self.sc = sc
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
        cam = sc.cam
        self.view, self.proj, self.campos, self.bg = t(cam.viewmatrix), t(cam.projmatrix), t(cam.campos), t(cam.bg)
        g = torch.Generator(device="cpu").manual_seed(seed + 1)
        self.anchor = torch.nn.Parameter(t(sc.means3D))
        base = torch.log(t(sc.scales).mean(dim=1, keepdim=True) * 2.0)
        self.scaling = torch.nn.Parameter(base.repeat(1, 6).contiguous())            # (N, 6) log-scales: offsets | cov
        self.rotation = torch.nn.Parameter(t(sc.rotations), requires_grad=False)
        self.offset = torch.nn.Parameter((torch.randn((n_anchor, k, 3), generator=g) * 0.5).to(device))
        self.anchor_feat = torch.nn.Parameter((torch.randn((n_anchor, feat_dim), generator=g) * 0.5).to(device))
        torch.manual_seed(seed + 2)
        mlp = lambda o, act: torch.nn.Sequential(torch.nn.Linear(feat_dim + 3, feat_dim), torch.nn.ReLU(True),  # noqa: E731
                                                 torch.nn.Linear(feat_dim, o), act).to(device)
        self.mlp_opacity, self.mlp_color, self.mlp_cov = mlp(k, torch.nn.Tanh()), mlp(3 * k, torch.nn.Sigmoid()), mlp(7 * k, torch.nn.Identity())
        params = [self.anchor, self.scaling, self.offset, self.anchor_feat] + [p for m in (self.mlp_opacity, self.mlp_color, self.mlp_cov) for p in m.parameters()]
        self.optimizer = torch.optim.Adam(params, lr=1e-3, eps=1e-15)
        self.opacity_accum = torch.zeros((n_anchor, 1), device=device)
        self.anchor_demon = torch.zeros((n_anchor, 1), device=device)
        self.offset_gradient_accum = torch.zeros((n_anchor * k, 1), device=device)
        self.offset_denom = torch.zeros((n_anchor * k, 1), device=device)
        yy, xx = torch.meshgrid(torch.arange(H, device=device).float() / H, torch.arange(W, device=device).float() / W, indexing="ij")
        self.gt = torch.stack([0.5 + 0.4 * torch.sin(6 * xx + 2 * yy), 0.5 + 0.4 * torch.cos(5 * yy), 0.5 + 0.4 * torch.sin(4 * (xx - yy))])
        g1 = torch.tensor([math.exp(-(x - 5) ** 2 / (2 * 1.5 ** 2)) for x in range(11)])
        g1 = (g1 / g1.sum()).unsqueeze(1)
        self.window = g1.mm(g1.t()).float()[None, None].expand(3, 1, 11, 11).contiguous().to(device)
        from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
        import scaffold_filter
        self.Settings, self.Rasterizer, self.filter_mod = GaussianRasterizationSettings, GaussianRasterizer, scaffold_filter
        if impl != "ours":
            from oracle.refcuda import RefSurfel
            self.ref = RefSurfel()

    def prefilter_voxel(self):
        scales = torch.exp(self.scaling)
        if self.impl == "ours":
            fs = self.filter_mod.GaussianRasterizationSettings(self.H, self.W, self.sc.cam.tanfovx, self.sc.cam.tanfovy, self.bg, 1.0,
                                                               self.view, self.proj, 0, self.campos, False, False)
            radii = self.filter_mod.GaussianRasterizer(fs).visible_filter(means3D=self.anchor, scales=scales[:, :3],
                                                                          rotations=self.rotation, cov3D_precomp=None)
        else:
            from oracle.refcuda import ref_visible_filter
            radii = ref_visible_filter(self.anchor.detach(), scales[:, :3].detach().contiguous(), self.rotation, self.view, self.proj,
                                       self.W, self.H, self.sc.cam.tanfovx, self.sc.cam.tanfovy)
        return radii > 0

    def generate_neural_gaussians(self, visible_mask):
        k = self.k
        feat, anchor = self.anchor_feat[visible_mask], self.anchor[visible_mask]
        grid_offsets, grid_scaling = self.offset[visible_mask], torch.exp(self.scaling)[visible_mask]
        ob_view = anchor - self.campos
        ob_view = ob_view / ob_view.norm(dim=1, keepdim=True)
        x = torch.cat([feat, ob_view], dim=1)
        neural_opacity = self.mlp_opacity(x).reshape([-1, 1])
        mask = (neural_opacity > 0.0).view(-1)
        opacity = neural_opacity[mask]
        color = self.mlp_color(x).reshape([anchor.shape[0] * k, 3])
        scale_rot = self.mlp_cov(x).reshape([anchor.shape[0] * k, 7])
        offsets = grid_offsets.view([-1, 3])
        rep = torch.cat([grid_scaling, anchor], dim=-1).repeat_interleave(k, dim=0)
        masked = torch.cat([rep, color, scale_rot, offsets], dim=-1)[mask]
        scaling_repeat, repeat_anchor, color, scale_rot, offsets = masked.split([6, 3, 3, 7, 3], dim=-1)
        scaling = scaling_repeat[:, 3:] * torch.sigmoid(scale_rot[:, :3])
        rot = F.normalize(scale_rot[:, 3:7])
        xyz = repeat_anchor + offsets * scaling_repeat[:, :3]
        return xyz, color, opacity, scaling, rot, neural_opacity, mask

    def step(self):
        visible = self.prefilter_voxel()
        xyz, color, opacity, scaling3, rot, neural_opacity, mask = self.generate_neural_gaussians(visible)
        scaling = scaling3[:, :2]                                   # stride-3 view, as GS-SR passes it
        screenspace_points = torch.zeros_like(xyz, requires_grad=True) + 0
        screenspace_points.retain_grad()
        rs = self.Settings(image_height=self.H, image_width=self.W, tanfovx=self.sc.cam.tanfovx, tanfovy=self.sc.cam.tanfovy,
                           bg=self.bg, scale_modifier=1.0, viewmatrix=self.view, projmatrix=self.proj, sh_degree=0,
                           campos=self.campos, prefiltered=False, debug=False)
        if self.impl == "ours":
            image, radii, allmap = self.Rasterizer(raster_settings=rs)(
                means3D=xyz, means2D=screenspace_points, shs=None, colors_precomp=color, opacities=opacity,
                scales=scaling, rotations=rot, cov3D_precomp=None)
        else:
            image, radii, allmap = _RefSurfelFn.apply(xyz, screenspace_points, None, color, opacity, scaling, rot, self.ref, rs)
        if self.fused_post:
            from gsr_b200.surfel_post import surfel_postprocess
            out = {"render": image, **surfel_postprocess(allmap, self.view, self.proj, depth_ratio=0.0)}
        else:
            render_alpha = allmap[1:2]
            render_normal = (allmap[2:5].permute(1, 2, 0) @ (self.view[:3, :3].T)).permute(2, 0, 1)
            depth = torch.nan_to_num(allmap[0:1] / render_alpha, 0, 0)
            surf_normal = self.depth_to_normal(depth).permute(2, 0, 1) * render_alpha.detach()
            out = {"render": image, "normal": render_normal, "surf_normal": surf_normal, "rend_dist": allmap[6:7]}
        losses = self.loss_dict(out)
        losses["scaling_loss"] = self.lambda_scaling * scaling.prod(dim=1).mean()
        loss = sum(losses.values())
        loss.backward()
        loss_value = loss.item()
        with torch.no_grad():                                       # scaffold_gaussian.py:488-508
            update_filter = radii > 0
            temp_opacity = neural_opacity.clone().view(-1).detach()
            temp_opacity[temp_opacity < 0] = 0
            self.opacity_accum[visible] += temp_opacity.view([-1, self.k]).sum(dim=1, keepdim=True)
            self.anchor_demon[visible] += 1
            av = visible.unsqueeze(dim=1).repeat([1, self.k]).view(-1)
            combined = torch.zeros_like(self.offset_gradient_accum, dtype=torch.bool).squeeze(dim=1)
            combined[av] = mask
            tmp = combined.clone()
            combined[tmp] = update_filter
            grad_norm = torch.norm(screenspace_points.grad[update_filter, :2], dim=-1, keepdim=True)
            self.offset_gradient_accum[combined] += grad_norm
            self.offset_denom[combined] += 1
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=True)
        self.last = dict(visible=int(visible.sum()), gaussians=int(mask.sum()), rendered=int(update_filter.sum()))
        return loss_value, {k_: float(v.detach()) for k_, v in losses.items()}


class _RefPlaneFn(torch.autograd.Function):
    """Reference diff-plane-rasterization kernels as an autograd op (L/diff_plane_rasterization/__init__.py:48-171)."""

    @staticmethod
    def forward(ctx, means3D, means2D, means2D_abs, colors, opacities, scales, rotations, all_map, ref, rs):
        o = ref.forward(rs.bg, rs.viewmatrix, rs.projmatrix, rs.campos, rs.image_width, rs.image_height, rs.tanfovx, rs.tanfovy,
                        means3D.contiguous(), opacities.contiguous(), scales.contiguous(), rotations.contiguous(),
                        colors=colors.contiguous(), all_map=all_map.contiguous(), render_geo=True)
        ctx.ref = ref
        ctx.mark_non_differentiable(o["radii"], o["observe"])
        return o["color"], o["radii"], o["observe"], o["out_all_map"], o["plane_depth"]

    @staticmethod
    def backward(ctx, g_color, g_radii, g_obs, g_all_map, g_plane_depth):
        g = ctx.ref.backward(g_color, g_all_map, g_plane_depth)
        return (g["means3D"], g["means2D"], g["means2D_abs"], g["colors"], g["opacities"], g["scales"], g["rotations"],
                g["all_map"], None, None)


class MiniPGSRTrainer(MiniTwoDGSTrainer):
    """PGSR iteration (BASELINE config 4): the reference view AND one neighbour view are rendered through the
    plane rasterizer (pgsr_scene.py:206-224), per-Gaussian all_map = [n_view, 1, |n_view . p_view|] built with autograd
    from the smallest-scale axis (pgsr_scene.py:238-256, 295-302), means2D / means2D_abs protocol (:262-268), depth normal
    from plane_depth (graphics_utils.py:110-146), L1 + SSIM + single-view normal loss (pgsr_scene.py:100-113, without the
    image-gradient weight) + a cross-view plane-depth term standing in for the multi-view geometric loss, densification
    statistics incl. the abs accumulator and out_observe gating (pgsr_gaussian.py:157-171)."""

    def __init__(self, P=100_000, W=800, H=450, seed=0, impl="ours", device="cuda"):
        self.impl, self.W, self.H, self.device = impl, W, H, device
        self.lambda_dssim, self.lambda_normal = 0.2, 0.015
        sc = synth.make_scene(P, W, H, seed=seed, scale_dims=3)
        self.sc = sc
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
        q = np.array([1.0, 0.004, -0.006, 0.002]); q /= np.linalg.norm(q)
        cams = [sc.cam, synth.make_camera(W, H, R=synth.quat_to_rot(q), t=np.array([0.06, -0.02, 0.03]))]
        self.cams = [dict(view=t(c.viewmatrix), proj=t(c.projmatrix), campos=t(c.campos), tanfovx=c.tanfovx, tanfovy=c.tanfovy) for c in cams]
        self.bg = t(sc.cam.bg)
        self.xyz = torch.nn.Parameter(t(sc.means3D))
        self.colors_raw = torch.nn.Parameter(torch.logit(t(sc.colors).clamp(1e-3, 1 - 1e-3)))
        self.scaling = torch.nn.Parameter(torch.log(t(sc.scales)))
        self.rotation = torch.nn.Parameter(t(sc.rotations))
        op = t(sc.opacities).clamp(1e-4, 1 - 1e-4)
        self.opacity = torch.nn.Parameter(torch.log(op / (1 - op)))
        self.optimizer = torch.optim.Adam([self.xyz, self.colors_raw, self.scaling, self.rotation, self.opacity], lr=1e-3, eps=1e-15)
        z = lambda *sh: torch.zeros(sh, device=device)  # noqa: E731
        self.xyz_gradient_accum, self.xyz_gradient_accum_abs, self.denom, self.max_radii2D = z(P, 1), z(P, 1), z(P, 1), z(P)
        yy, xx = torch.meshgrid(torch.arange(H, device=device).float() / H, torch.arange(W, device=device).float() / W, indexing="ij")
        self.gt = torch.stack([0.5 + 0.4 * torch.sin(6 * xx + 2 * yy), 0.5 + 0.4 * torch.cos(5 * yy), 0.5 + 0.4 * torch.sin(4 * (xx - yy))])
        g1 = torch.tensor([math.exp(-(x - 5) ** 2 / (2 * 1.5 ** 2)) for x in range(11)])
        g1 = (g1 / g1.sum()).unsqueeze(1)
        self.window = g1.mm(g1.t()).float()[None, None].expand(3, 1, 11, 11).contiguous().to(device)
        from diff_plane_rasterization import GaussianRasterizationSettings, GaussianRasterizer
        self.Settings, self.Rasterizer = GaussianRasterizationSettings, GaussianRasterizer
        if impl != "ours":
            from oracle.refcuda import RefGauss
            self.refs = [RefGauss(plane=True), RefGauss(plane=True)]     # one context per view kept alive until backward

    @staticmethod
    def quaternion_to_matrix(q):                                        # pytorch3d.transforms.quaternion_to_matrix
        r, i, j, k = torch.unbind(q, -1)
        two_s = 2.0 / (q * q).sum(-1)
        o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                         two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                         two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
        return o.reshape(q.shape[:-1] + (3, 3))

    def normal_from_depth(self, depth, cam):                            # graphics_utils.py:80-146 (offset=None)
        H, W = depth.shape
        fx, fy = W / (2 * cam["tanfovx"]), H / (2 * cam["tanfovy"])
        ix, iy = torch.meshgrid(torch.arange(W, device=self.device).float(), torch.arange(H, device=self.device).float(), indexing="xy")
        xyz = torch.stack([(ix - W / 2) / fx * depth, (iy - H / 2) / fy * depth, depth], -1)
        l2r = xyz[1:H - 1, 2:W] - xyz[1:H - 1, 0:W - 2]
        b2t = xyz[0:H - 2, 1:W - 1] - xyz[2:H, 1:W - 1]
        n = F.normalize(torch.cross(l2r, b2t, dim=-1), p=2, dim=-1)
        return F.pad(n.permute(2, 0, 1), (1, 1, 1, 1), mode="constant")

    def gaussians_for_view(self, vi):
        return (self.xyz, torch.sigmoid(self.opacity), torch.exp(self.scaling), F.normalize(self.rotation),
                torch.sigmoid(self.colors_raw))

    def render_view(self, vi):
        cam = self.cams[vi]
        means3D, opacity, scales, rotations, colors = self.gaussians_for_view(vi)
        sp = torch.zeros_like(means3D, requires_grad=True) + 0
        sp_abs = torch.zeros_like(means3D, requires_grad=True) + 0
        sp.retain_grad(); sp_abs.retain_grad()
        rs = self.Settings(image_height=self.H, image_width=self.W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=self.bg,
                           scale_modifier=1.0, viewmatrix=cam["view"], projmatrix=cam["proj"], sh_degree=0, campos=cam["campos"],
                           prefiltered=False, render_geo=True, debug=False)
        Rm = self.quaternion_to_matrix(rotations)
        idx = scales.min(dim=-1)[1][..., None, None].expand(-1, 3, -1)
        normal_global = Rm.gather(2, idx).squeeze(dim=2)
        neg = (normal_global * (cam["campos"] - means3D)).sum(-1) < 0.0
        normal_global = torch.where(neg[:, None], -normal_global, normal_global)
        local_normal = normal_global @ cam["view"][:3, :3]
        pts_in_cam = means3D @ cam["view"][:3, :3] + cam["view"][3, :3]
        local_distance = (local_normal * pts_in_cam).sum(-1).abs()
        all_map = torch.cat([local_normal, torch.ones_like(local_distance)[:, None], local_distance[:, None]], dim=1)
        if self.impl == "ours":
            image, radii, observe, out_all_map, plane_depth = self.Rasterizer(raster_settings=rs)(
                means3D=means3D, means2D=sp, means2D_abs=sp_abs, shs=None, colors_precomp=colors, opacities=opacity,
                scales=scales, rotations=rotations, all_map=all_map, cov3D_precomp=None)
        else:
            image, radii, observe, out_all_map, plane_depth = _RefPlaneFn.apply(means3D, sp, sp_abs, colors, opacity, scales,
                                                                                rotations, all_map, self.refs[vi], rs)
        rendered_normal, rendered_alpha = out_all_map[0:3], out_all_map[3:4]
        if getattr(self, "fused_post", False):      # fused depth -> normal x alpha (gsr_b200.depth_normal, csrc/depth_normal.cu)
            from gsr_b200.depth_normal import render_normal_weighted
            fx, fy = self.W / (2 * cam["tanfovx"]), self.H / (2 * cam["tanfovy"])
            K = torch.tensor([[fx, 0, self.W / 2], [0, fy, self.H / 2], [0, 0, 1]], dtype=torch.float32)   # get_calib_matrix_nerf
            depth_normal = render_normal_weighted(plane_depth.squeeze(), K, rendered_alpha.squeeze(0))
        else:
            depth_normal = self.normal_from_depth(plane_depth.squeeze(), cam) * rendered_alpha.detach()
        return dict(render=image, viewspace_points=sp, viewspace_points_abs=sp_abs, visibility_filter=radii > 0, radii=radii,
                    out_observe=observe, rendered_normal=rendered_normal, plane_depth=plane_depth,
                    rendered_distance=out_all_map[4:5], depth_normal=depth_normal, scaling=scales)

    def step(self):
        ref, near = self.render_view(0), self.render_view(1)
        image = ref["render"]
        losses = {"L1_loss": (1.0 - self.lambda_dssim) * torch.abs(image - self.gt).mean(),
                  "ssim_loss": self.lambda_dssim * (1.0 - self.ssim(image, self.gt)),
                  "normal_loss": self.lambda_normal * ((ref["depth_normal"] - ref["rendered_normal"]).abs().sum(0)).mean(),
                  "geo_loss": 0.03 * (ref["plane_depth"] - near["plane_depth"]).abs().clamp(max=1.0).mean()
                              + 0.01 * (ref["rendered_distance"] - near["rendered_distance"]).abs().clamp(max=1.0).mean()}
        loss = sum(losses.values())
        loss.backward()
        loss_value = loss.item()
        with torch.no_grad():                                           # pgsr_gaussian.py:157-171
            vis, radii = ref["visibility_filter"], ref["radii"]
            m = (ref["out_observe"] > 0) & vis
            self.max_radii2D[m] = torch.max(self.max_radii2D[m], radii[m].float())
            self.xyz_gradient_accum[vis] += torch.norm(ref["viewspace_points"].grad[vis, :2], dim=-1, keepdim=True)
            self.xyz_gradient_accum_abs[vis] += torch.norm(ref["viewspace_points_abs"].grad[vis, :2], dim=-1, keepdim=True)
            self.denom[vis] += 1
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=True)
        self.last = dict(observed=int((ref["out_observe"] > 0).sum()), visible=int(vis.sum()))
        return loss_value, {k_: float(v.detach()) for k_, v in losses.items()}


# ---- octree variants (BASELINE configs 4 and 5) and anchor growing ------------------------------------------------------
class OctreeMixin:
    """Level-of-detail anchors of OctreeGaussian, restated: every anchor carries an integer `level`; per view,
    set_anchor_mask keeps the anchors whose level does not exceed the level predicted from the camera distance
    (gssr/gaussian/octree_gaussian.py:184-196 dist2level='round', :255-267), prefilter_voxel runs visible_filter on
    the masked anchors only and scatters the result back (gssr/scene/octree_scene.py:136-172), and the MLPs see the level
    as an extra input (add_level, octree_scene.py:42-66)."""
    fork, levels = 2, 4

    def init_octree(self, n_anchor, seed):
        g = torch.Generator(device="cpu").manual_seed(seed + 7)
        self.level = torch.randint(0, self.levels, (n_anchor, 1), generator=g).to(self.device)
        self.extra_level = torch.zeros(n_anchor, device=self.device)
        self.voxel_size = 0.01
        d = (self.anchor.detach() - self.campos).norm(dim=1)
        self.standard_dist = float(d.quantile(0.25))                     # nearer anchors may use the finest levels
        self.anchor_mask = torch.ones(n_anchor, dtype=torch.bool, device=self.device)

    def set_anchor_mask(self, cam_center, resolution_scale=1.0):
        anchor_pos = self.anchor + (self.voxel_size / 2) / (float(self.fork) ** self.level)
        dist = torch.sqrt(torch.sum((anchor_pos - cam_center) ** 2, dim=1)) * resolution_scale
        pred_level = torch.log2(self.standard_dist / dist) / math.log2(self.fork) + self.extra_level
        int_level = torch.clamp(torch.round(pred_level).int(), min=0, max=self.levels - 1)
        self.anchor_mask = (self.level.squeeze(dim=1) <= int_level)

    def octree_prefilter(self, view, proj, campos, tanfovx, tanfovy):
        anchor_mask = self.anchor_mask
        means3D = self.anchor[anchor_mask]
        scales = torch.exp(self.scaling)[anchor_mask]
        rotations = self.rotation[anchor_mask]
        if self.impl == "ours":
            fs = self.filter_mod.GaussianRasterizationSettings(self.H, self.W, tanfovx, tanfovy, self.bg, 1.0, view, proj, 0, campos,
                                                               False, False)
            radii_pure = self.filter_mod.GaussianRasterizer(fs).visible_filter(means3D=means3D, scales=scales[:, :3],
                                                                               rotations=rotations, cov3D_precomp=None)
        else:
            from oracle.refcuda import ref_visible_filter
            radii_pure = ref_visible_filter(means3D.detach().contiguous(), scales[:, :3].detach().contiguous(), rotations.contiguous(),
                                            view, proj, self.W, self.H, tanfovx, tanfovy)
        visible_mask = anchor_mask.clone()
        visible_mask[anchor_mask] = radii_pure > 0
        return visible_mask


class MiniOctree2DGSTrainer(OctreeMixin, MiniScaffold2DGSTrainer):
    """Octree-2DGS iteration (BASELINE config 5's per-tile model; octree_2dgs_scene.py:11-27 = TwoDGSScene + OctreeScene):
    set_anchor_mask -> octree prefilter_voxel -> neural Gaussians with the level input -> surfel rasterizer on the
    stride-3 scaling[:, :2] view -> 2DGS + scaling losses -> anchor statistics."""

    def __init__(self, n_anchor=200_000, k=5, feat_dim=32, **kw):
        super().__init__(n_anchor=n_anchor, k=k, feat_dim=feat_dim, **kw)
        self.init_octree(n_anchor, kw.get("seed", 0))
        torch.manual_seed(kw.get("seed", 0) + 3)
        mlp = lambda o, act: torch.nn.Sequential(torch.nn.Linear(feat_dim + 3 + 1, feat_dim), torch.nn.ReLU(True),  # noqa: E731
                                                 torch.nn.Linear(feat_dim, o), act).to(self.device)
        self.mlp_opacity, self.mlp_color, self.mlp_cov = mlp(k, torch.nn.Tanh()), mlp(3 * k, torch.nn.Sigmoid()), mlp(7 * k, torch.nn.Identity())
        params = [self.anchor, self.scaling, self.offset, self.anchor_feat] + [p for m in (self.mlp_opacity, self.mlp_color, self.mlp_cov) for p in m.parameters()]
        self.optimizer = torch.optim.Adam(params, lr=1e-3, eps=1e-15)

    def prefilter_voxel(self):
        self.set_anchor_mask(self.campos)
        return self.octree_prefilter(self.view, self.proj, self.campos, self.sc.cam.tanfovx, self.sc.cam.tanfovy)

    def generate_neural_gaussians(self, visible_mask):
        k = self.k
        feat, anchor, level = self.anchor_feat[visible_mask], self.anchor[visible_mask], self.level[visible_mask].float()
        grid_offsets, grid_scaling = self.offset[visible_mask], torch.exp(self.scaling)[visible_mask]
        ob_view = anchor - self.campos
        ob_view = ob_view / ob_view.norm(dim=1, keepdim=True)
        x = torch.cat([feat, ob_view, level], dim=1)
        neural_opacity = self.mlp_opacity(x).reshape([-1, 1])
        mask = (neural_opacity > 0.0).view(-1)
        opacity = neural_opacity[mask]
        color = self.mlp_color(x).reshape([anchor.shape[0] * k, 3])
        scale_rot = self.mlp_cov(x).reshape([anchor.shape[0] * k, 7])
        rep = torch.cat([grid_scaling, anchor], dim=-1).repeat_interleave(k, dim=0)
        masked = torch.cat([rep, color, scale_rot, grid_offsets.view([-1, 3])], dim=-1)[mask]
        scaling_repeat, repeat_anchor, color, scale_rot, offsets = masked.split([6, 3, 3, 7, 3], dim=-1)
        scaling = scaling_repeat[:, 3:] * torch.sigmoid(scale_rot[:, :3])
        rot = F.normalize(scale_rot[:, 3:7])
        xyz = repeat_anchor + offsets * scaling_repeat[:, :3]
        return xyz, color, opacity, scaling, rot, neural_opacity, mask


class MiniOctreePGSRTrainer(OctreeMixin, MiniPGSRTrainer):
    """Octree-PGSR iteration (BASELINE config 4; octree_pgsr_scene.py:13-47 = PGSRScene + OctreeScene): per view
    set_anchor_mask + octree prefilter + neural Gaussians, rendered through the plane rasterizer with the autograd-built
    all_map; PGSR losses + scaling loss; anchor statistics from the reference view."""

    def __init__(self, n_anchor=100_000, k=5, feat_dim=32, W=800, H=450, seed=0, impl="ours", device="cuda"):
        super().__init__(P=n_anchor, W=W, H=H, seed=seed, impl=impl, device=device)
        self.k = k
        sc = self.sc
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
        g = torch.Generator(device="cpu").manual_seed(seed + 1)
        self.anchor = torch.nn.Parameter(t(sc.means3D))
        base = torch.log(t(sc.scales).mean(dim=1, keepdim=True) * 2.0)
        self.scaling = torch.nn.Parameter(base.repeat(1, 6).contiguous())
        self.rotation = torch.nn.Parameter(t(sc.rotations), requires_grad=False)
        self.offset = torch.nn.Parameter((torch.randn((n_anchor, k, 3), generator=g) * 0.5).to(device))
        self.anchor_feat = torch.nn.Parameter((torch.randn((n_anchor, feat_dim), generator=g) * 0.5).to(device))
        self.campos = self.cams[0]["campos"]
        self.init_octree(n_anchor, seed)
        torch.manual_seed(seed + 2)
        mlp = lambda o, act: torch.nn.Sequential(torch.nn.Linear(feat_dim + 3 + 1, feat_dim), torch.nn.ReLU(True),  # noqa: E731
                                                 torch.nn.Linear(feat_dim, o), act).to(device)
        self.mlp_opacity, self.mlp_color, self.mlp_cov = mlp(k, torch.nn.Tanh()), mlp(3 * k, torch.nn.Sigmoid()), mlp(7 * k, torch.nn.Identity())
        params = [self.anchor, self.scaling, self.offset, self.anchor_feat] + [p for m in (self.mlp_opacity, self.mlp_color, self.mlp_cov) for p in m.parameters()]
        self.optimizer = torch.optim.Adam(params, lr=1e-3, eps=1e-15)
        self.opacity_accum = torch.zeros((n_anchor, 1), device=device)
        self.anchor_demon = torch.zeros((n_anchor, 1), device=device)
        self.offset_gradient_accum = torch.zeros((n_anchor * k, 1), device=device)
        self.offset_denom = torch.zeros((n_anchor * k, 1), device=device)
        import scaffold_filter
        self.filter_mod = scaffold_filter
        self.lambda_scaling = 0.01

    generate_neural_gaussians = MiniOctree2DGSTrainer.generate_neural_gaussians

    def gaussians_for_view(self, vi):
        cam = self.cams[vi]
        self.campos = cam["campos"]
        self.set_anchor_mask(cam["campos"])
        visible = self.octree_prefilter(cam["view"], cam["proj"], cam["campos"], cam["tanfovx"], cam["tanfovy"])
        xyz, color, opacity, scaling, rot, neural_opacity, mask = self.generate_neural_gaussians(visible)
        self._stat = (visible, neural_opacity, mask) if vi == 0 else self._stat
        return xyz, opacity, scaling, rot, color

    def step(self):
        ref, near = self.render_view(0), self.render_view(1)
        image = ref["render"]
        losses = {"L1_loss": (1.0 - self.lambda_dssim) * torch.abs(image - self.gt).mean(),
                  "ssim_loss": self.lambda_dssim * (1.0 - self.ssim(image, self.gt)),
                  "normal_loss": self.lambda_normal * ((ref["depth_normal"] - ref["rendered_normal"]).abs().sum(0)).mean(),
                  "geo_loss": 0.03 * (ref["plane_depth"] - near["plane_depth"]).abs().clamp(max=1.0).mean(),
                  "scaling_loss": self.lambda_scaling * ref["scaling"].prod(dim=1).mean()}
        loss = sum(losses.values())
        loss.backward()
        loss_value = loss.item()
        with torch.no_grad():                                           # scaffold_gaussian.py:488-508 on the reference view
            visible, neural_opacity, mask = self._stat
            update_filter = ref["radii"] > 0
            temp_opacity = neural_opacity.clone().view(-1).detach()
            temp_opacity[temp_opacity < 0] = 0
            self.opacity_accum[visible] += temp_opacity.view([-1, self.k]).sum(dim=1, keepdim=True)
            self.anchor_demon[visible] += 1
            av = visible.unsqueeze(dim=1).repeat([1, self.k]).view(-1)
            combined = torch.zeros_like(self.offset_gradient_accum, dtype=torch.bool).squeeze(dim=1)
            combined[av] = mask
            tmp = combined.clone()
            combined[tmp] = update_filter
            self.offset_gradient_accum[combined] += torch.norm(ref["viewspace_points"].grad[update_filter, :2], dim=-1, keepdim=True)
            self.offset_denom[combined] += 1
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=True)
        self.last = dict(observed=int((ref["out_observe"] > 0).sum()), visible=int(visible.sum()), gaussians=int(mask.sum()))
        return loss_value, {k_: float(v.detach()) for k_, v in losses.items()}


def scatter_max_rows(src, index, n_out):
    """torch_scatter.scatter_max(src, index[:, None].expand_as(src), dim=0)[0] (scaffold_gaussian.py:619): per output row the
    column-wise maximum over the source rows mapped to it; rows nobody maps to stay 0."""
    out = torch.zeros((n_out, src.shape[1]), dtype=src.dtype, device=src.device)
    return out.scatter_reduce(0, index[:, None].expand_as(src), src, reduce="amax", include_self=False)


def anchor_growing_candidates(tr, grad_threshold=0.0002, update_depth=3, update_init_factor=16, update_hierachy_factor=4,
                              check_interval=100, success_threshold=0.8, seed=0):
    """The selection half of ScaffoldGaussian.adjust_anchor / anchor_growing (gssr/gaussian/scaffold_gaussian.py:557-640,
    :642-651): from the statistics the rasterizer's means2D gradients fed (offset_gradient_accum / offset_denom), the
    voxel-snapped positions and max-pooled features of the anchors that would be added at each of the `update_depth`
    levels.  Returns [(candidate_anchor (n,3), new_feat (n,F))]; deterministic given `seed` (the reference draws
    torch.rand_like masks).  The optimizer surgery that follows in the reference (cat_tensors_to_optimizer) is not part
    of what the rasterizer influences and is left out."""
    k = tr.k
    grads = tr.offset_gradient_accum / tr.offset_denom
    grads[grads.isnan()] = 0.0
    grads = torch.norm(grads, dim=-1)
    offset_mask = (tr.offset_denom > check_interval * success_threshold * 0.5).squeeze(dim=1)
    gen = torch.Generator(device=grads.device).manual_seed(seed)
    scaling = torch.exp(tr.scaling)
    out = []
    for i in range(update_depth):
        cur_threshold = grad_threshold * ((update_hierachy_factor // 2) ** i)
        candidate_mask = (grads >= cur_threshold) & offset_mask
        rand_mask = torch.rand(candidate_mask.shape, generator=gen, device=grads.device) > (0.5 ** (i + 1))
        candidate_mask = candidate_mask & rand_mask
        all_xyz = tr.anchor.unsqueeze(dim=1) + tr.offset * scaling[:, :3].unsqueeze(dim=1)
        cur_size = tr.voxel_size * (update_init_factor // (update_hierachy_factor ** i))
        grid_coords = torch.round(tr.anchor / cur_size).int()
        selected_grid_coords = torch.round(all_xyz.view([-1, 3])[candidate_mask] / cur_size).int()
        uniq, inverse = torch.unique(selected_grid_coords, return_inverse=True, dim=0)
        # drop voxels that already hold an anchor (hash compare instead of the reference's chunked broadcast compare)
        key = lambda c: (c[:, 0].long() * 73856093) ^ (c[:, 1].long() * 19349663) ^ (c[:, 2].long() * 83492791)  # noqa: E731
        occupied = torch.isin(key(uniq), key(grid_coords))
        exact = torch.zeros_like(occupied)
        if occupied.any():                                              # confirm hash hits exactly
            cand = uniq[occupied]
            exact_hit = (cand.unsqueeze(1) == grid_coords[torch.isin(key(grid_coords), key(cand))].unsqueeze(0)).all(-1).any(-1)
            exact[occupied] = exact_hit
        keep = ~exact
        candidate_anchor = uniq[keep] * cur_size
        new_feat = tr.anchor_feat.unsqueeze(dim=1).repeat([1, k, 1]).view([-1, tr.anchor_feat.shape[1]])[candidate_mask]
        new_feat = scatter_max_rows(new_feat, inverse, uniq.shape[0])[keep]
        out.append((candidate_anchor, new_feat))
    return out


def measure_graph_iters_per_s(P=100_000, W=800, H=800, iters=30, warmup=5, seed=0, fused_ssim=True, fused_post=True):
    """2DGS iterations replayed from ONE CUDA graph (drop-in rasterizer + fused image-space ops; the reference arm has no
    counterpart, its forward cannot be captured).  Returns (iters/s, last loss, capture_overflow())."""
    from gsr_b200.graphs import capture_overflow
    tr = MiniTwoDGSTrainer(P, W, H, seed=seed, impl="ours", capturable=True)
    tr.fused_ssim, tr.fused_post = fused_ssim, fused_post
    graph, loss = tr.capture(warmup=max(warmup, 3))
    capture_overflow(reset=True)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return iters / (e0.elapsed_time(e1) * 1e-3), float(loss), capture_overflow(reset=True)


def measure_iters_per_s(impl, P=100_000, W=800, H=800, iters=30, warmup=5, seed=0, scaffold=False, pgsr=False,
                        fused_ssim=False, fused_post=False, octree=False):
    if octree and pgsr:
        tr = MiniOctreePGSRTrainer(P, W=W, H=H, seed=seed, impl=impl)
    elif octree:
        tr = MiniOctree2DGSTrainer(P, W=W, H=H, seed=seed, impl=impl)
    elif pgsr:
        tr = MiniPGSRTrainer(P, W=W, H=H, seed=seed, impl=impl)
    elif scaffold:
        tr = MiniScaffold2DGSTrainer(P, W=W, H=H, seed=seed, impl=impl)
    else:
        tr = MiniTwoDGSTrainer(P, W, H, seed=seed, impl=impl)
    tr.fused_ssim, tr.fused_post = fused_ssim, fused_post
    for _ in range(warmup):
        tr.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        last = tr.step()
    e1.record()
    torch.cuda.synchronize()
    return iters / (e0.elapsed_time(e1) * 1e-3), last
