"""Forward pass of the Nature CNN entirely on the tcgen05 kernels (second slice).

`NatureCnnTcForward` takes the weights of a `NatureCNN` (torch layout: conv [Cout, Cin, KH, KW], NCHW flatten)
and re-lays them once for the implicit-GEMM convolution: NHWC kernels with K ordered (kh, kw, c), the two
strided layers rewritten as stride-1 layers over space-to-depth inputs (8x8/4 -> 2x2/1 on 21x21x64,
4x4/2 -> 2x2/1 on 10x10x128), the FC weight permuted from (c, h, w) to (h, w, c) flatten order (which is the
order the reference's Keras Flatten uses on NHWC), both heads stacked into one [8, 512] matrix.

    uint8 frames --xa_conv2d_u8_s2d_bf16 (/255, space-to-depth, 2x2 conv, ReLU, out_s2d: one kernel)--> [B,10,10,128]
      --xa_conv2d_nhwc_bf16 (2x2, ReLU)-->             [B,9,9,64]
      --xa_conv2d_nhwc_bf16 (3x3, ReLU)-->             [B,7,7,64] = [B,3136]
      --xa_gemm_bf16_tn (ReLU)--> [B,512] --xa_heads_forward_bf16--> logits [B,A], value [B]

Inference only (rollout-time policy evaluation and the bootstrap value: T+1 of the T+1+K*M forward passes of a
train step); training runs the same kernels plus the backward pass through `NatureCnnTc` (tc_cnn.py, tc_plan.py).
"""
import torch

from .. import ops
from .tc_operands import OperandPack


class NatureCnnTcForward:
    def __init__(self, module):
        self.module = module
        self._graphs = {}
        self._plans = {}                                          # batch size -> NaturePlan (kept: captured graphs read their buffers)
        self._pack, self._shared = None, False
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        """Re-derive the bf16 operand copies from the module's fp32 weights (after an optimiser step): two launches into
        buffers with fixed addresses, so captured graphs stay valid.  A `NatureCnnTc` module shares its own pack."""
        if self._pack is None:
            own = getattr(self.module, '_op', None) if hasattr(self.module, 'refresh') else None
            if own is None and hasattr(self.module, 'refresh'):
                own = self.module.refresh()._op
            self._pack = own if own is not None else OperandPack(self.module)
            self._shared = own is not None
        elif not self._shared:
            self._pack.refresh()
        for name in ('w1', 'b1', 'w2', 'b2', 'w3', 'b3', 'wf', 'bf_', 'wh', 'bh'):
            setattr(self, name, getattr(self._pack, name))
        self.n_actions = self._pack.n_actions
        return self

    @torch.no_grad()
    def graphed(self, batch):
        """A CUDA-graph replay of the six launches for a fixed batch size (the rollout's n_envs): at 256 frames the
        eager pipeline is bounded by host-side launch overhead, a replay is one launch.  Returns fn(frames_u8)."""
        if batch in self._graphs:
            return self._graphs[batch]
        dev = self.w1.device
        static_in = torch.zeros((batch, 84, 84, 4), dtype=torch.uint8, device=dev)
        stream = torch.cuda.Stream(dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream):
            self(static_in)                          # warm-up (module loading, allocator) outside the capture
            stream.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                static_out = self(static_in)
        torch.cuda.current_stream(dev).wait_stream(stream)

        def run(frames_u8):
            static_in.copy_(frames_u8)
            graph.replay()
            return static_out[0], static_out[1]

        self._graphs[batch] = run
        self._keepalive = getattr(self, '_keepalive', []) + [(graph, static_in, static_out)]
        return run

    @torch.no_grad()
    def __call__(self, frames_u8, out=None):
        """uint8 [B,84,84,4] -> (actor_out [B,A] fp32, critic [B] fp32): one native call, six launches (csrc/nature_net.cu; the
        first layer reads the frames directly).  The outputs are the plan's own buffers for this batch size, valid until the
        next call with it -- or the caller's `out` = (actor, critic) tensors (rows of the rollout buffers)."""
        batch = frames_u8.shape[0]
        plan = self._plans.get(batch)
        if plan is None:
            from .tc_plan import NaturePlan
            plan = self._plans[batch] = NaturePlan(self._pack, self._pack._named, None, batch, backward=False)
        return plan.forward(frames_u8 if frames_u8.is_contiguous() else frames_u8.contiguous(), out=out)
