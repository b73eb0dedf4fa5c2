"""CUDA-graph capture of the decoder forward.

One decoder forward is ~3000 small kernel launches around the sampling kernels; on B200 the GPU finishes them faster
than Python can issue them (profiles/r01_decoder_forward_kernel_breakdown.json: 25 ms of GPU work in a 39 ms step).
The whole forward is therefore captured once and replayed. libracformer_ops.so is capture-safe by construction (no
allocation, no synchronisation, no host reads; launches go to the stream it is given), and the harness creates no
host-side tensors during forward.
"""
import torch


class GraphedDecoderForward:
    """Capture `model(query_bbox, query_feat, mlvl_feats, lss_bev, radar_bev, None, meta)` and replay it.

    The tensors in `example` become the static input buffers: write new inputs into them (`load`) and call the
    object. Outputs are static too; clone them if they must survive the next replay.
    """

    def __init__(self, model, example, warmup=3):
        assert not model.training, "graph capture is for inference (dropout / checkpointing are host-driven)"
        dev = example["query_bbox"].device
        self.model = model
        self.static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in example.items() if k != "mlvl_feats"}
        self.static["mlvl_feats"] = [f.clone() for f in example["mlvl_feats"]]
        metas = example["img_metas"]
        self.meta = metas if isinstance(metas, dict) else model.decoder.build_meta(metas, example["query_bbox"].shape[0], dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from .caches import cache_epoch
        self.epoch = cache_epoch()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.outputs = self._run()

    def _run(self):
        s = self.static
        return self.model(s["query_bbox"], s["query_feat"], s["mlvl_feats"], s["lss_bev"], s["radar_bev"], None, self.meta)

    def load(self, inputs, non_blocking=True):
        """Copy new inputs (host or device tensors) into the static buffers on the current stream."""
        for k in ("query_bbox", "query_feat", "lss_bev", "radar_bev"):
            if k in inputs:
                self.static[k].copy_(inputs[k], non_blocking=non_blocking)
        if "mlvl_feats" in inputs:
            for dst, src in zip(self.static["mlvl_feats"], inputs["mlvl_feats"]):
                dst.copy_(src, non_blocking=non_blocking)

    def __call__(self, inputs=None):
        from .caches import cache_epoch
        if cache_epoch() != self.epoch:
            raise RuntimeError("GraphedDecoderForward: the weight caches were invalidated after capture (load_state_dict / "
                               "train() / invalidate_weight_caches()); the graph holds pointers to the old operand splits -- "
                               "capture a new one")
        if inputs is not None:
            self.load(inputs)
        self.graph.replay()
        return self.outputs


class PipelinedDecoderForward:
    """End-to-end serving loop for host-resident inputs: H2D copy of sample s+1 overlaps the decoder forward of sample s.

    `depth` captured graphs, each with its own static input/output buffers, are used round-robin. A copy stream
    moves pinned host inputs into the next slot while the compute stream replays the current one; results are copied
    back into pinned host buffers on the compute stream. Per sample: submit() -> ticket, then result(ticket).
    """

    def __init__(self, model, example, depth=2):
        self.dev = example["query_bbox"].device
        self.slots = [GraphedDecoderForward(model, example) for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.loaded = [torch.cuda.Event() for _ in range(depth)]     # H2D of the slot finished
        self.free = [torch.cuda.Event() for _ in range(depth)]       # compute + D2H of the slot finished
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.host_out = [[torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in s.outputs] for s in self.slots]
        for e in self.free:
            e.record(torch.cuda.current_stream(self.dev))
        self.next = 0

    def submit(self, host_inputs):
        i = self.next
        self.next = (self.next + 1) % len(self.slots)
        slot = self.slots[i]
        compute = torch.cuda.current_stream(self.dev)
        self.copy_stream.wait_event(self.free[i])                    # slot's previous sample fully consumed
        with torch.cuda.stream(self.copy_stream):
            slot.load(host_inputs, non_blocking=True)
            self.loaded[i].record(self.copy_stream)
        compute.wait_event(self.loaded[i])
        outs = slot()
        for h, o in zip(self.host_out[i], outs):
            h.copy_(o, non_blocking=True)
        self.free[i].record(compute)
        self.done[i].record(compute)
        return i

    def result(self, ticket):
        """The sample's outputs as fresh host tensors (the slot's pinned buffers are overwritten `depth` submits later)."""
        self.done[ticket].synchronize()
        return [h.clone() for h in self.host_out[ticket]]
