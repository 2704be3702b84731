"""Inference as the reference's front-ends run it (/root/reference/src/margipose/bin/infer_single.py:58-66,
bin/eval_3d.py:48-62): `model.eval()`, one forward per batch, 3D joint coordinates out.

    infer = InferStep(model, batch=1)
    xyz = infer(images)          # fp32 (B, 3, H, W) normalised, or uint8 (B, H, W, 3) raw pixels

Eval-mode BatchNorm is folded into the conv epilogues (engine.py `conv_folded`), all buffers are static, and
after `warmup` eager calls the whole forward -- ~370 launches for the 4-stage ResNet-34 model -- is captured
once into a CUDA graph: per call the host does one async copy and one graph launch.  With uint8 input the
reference's input step (ImageSpecs.convert: /255, ImageNet mean / stddev, data_specs.py:6-13,38-39) runs
inside the stem conv's gather kernel.  The multi-crop evaluation of eval_3d.py:67-79 is a batch of 10 crops
through this same call (the per-crop de-normalisation and averaging stay on the CPU in fp64, as there).
"""
import torch

from . import parallel  # noqa: F401  (keeps import order stable for torch.distributed users)


class InferStep:
    def __init__(self, model, batch, height=256, width=256, use_graph=True, warmup=2, uint8=False):
        dev = next(model.parameters()).device
        if dev.type != 'cuda':
            raise ValueError('InferStep needs the model on a CUDA device')
        self.model, self.device = model, dev
        model.eval()
        model._ensure(dev)
        self.eng = model.engine_for(batch, height, width, False)
        self.uint8 = bool(uint8)
        if self.uint8:
            self.x = torch.zeros(batch, height, width, 3, dtype=torch.uint8, device=dev)
            self.eng.x_u8 = self.x
        else:
            self.x = self.eng.x_in
        self.coords = None
        self.use_graph, self.warmup = use_graph, warmup
        self._graph = None
        self._eager_runs = 0

    def _forward(self):
        model = self.model
        probs = self.eng.forward(self.x)
        model.xy_heatmaps = [row[0] for row in probs]
        model.zy_heatmaps = [row[1] for row in probs]
        model.xz_heatmaps = [row[2] for row in probs]
        return model.heatmaps_to_coords(*probs[-1])

    @torch.no_grad()
    def run(self):
        """One forward on whatever is in the static input buffer `self.x`; returns the static (B, J, 3) output."""
        model = self.model
        if model.training:
            raise RuntimeError('InferStep runs the model in eval mode; call model.eval()')
        model._refresh_packs()      # outside the graph: repacks only after a parameter write
        if self.use_graph and self._graph is None and self._eager_runs >= self.warmup:
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.coords = self._forward()
            self._graph = g
        if self._graph is not None:
            self._graph.replay()
        else:
            self.coords = self._forward()
            self._eager_runs += 1
        return self.coords

    def __call__(self, images):
        self.x.copy_(images, non_blocking=True)
        return self.run()

    def launches(self):
        return self.eng.launches(self.eng.fwd) + 2      # + BatchNorm fold table, + coordinates
