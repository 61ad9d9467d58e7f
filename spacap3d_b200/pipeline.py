"""Throughput runner for the detector forward: CUDA-graph replay on several streams.

The forward is a chain of ~120 short, shape-static launches dominated by a latency-bound kernel
(FPS: 2047 sequential rounds on 64 of the 148 SMs).  Two things follow:
  * launch overhead is removed by capturing the whole forward once per stream into a CUDA graph
    (our C-ABI launches take the stream explicitly, so they are captured like any torch op);
  * consecutive batches are independent, so replaying them round-robin on 2-3 streams lets the FPS
    of batch k+1 run on idle SMs while the tensor-core / gather kernels of batch k run
    ("overlap ... on separate streams; capture launch-bound inner loops in CUDA graphs").
Results are bit-identical to the eager forward (same kernels, same order per batch).
"""
import torch


class GraphedDetector:
    """Round-robin CUDA-graph replay of `model({"point_clouds": x})` for a fixed input shape.

    submit(x) copies x (host pinned or device) into the slot's static input on the slot's stream,
    replays the graph and (optionally) copies the requested outputs to pinned host buffers; it
    returns the slot index.  Outputs of a slot stay valid until that slot is submitted again."""

    def __init__(self, model, example, n_streams=2, result_keys=None, warmup=3, sa_min_tiles=16, fps_algo=None,
                 pm_n_tile=256, pm_tiles_per_cta=4):
        assert example.is_cuda
        # Launch hints baked into the captured graphs (per call and thread-local: nothing process-wide changes).
        # With several batches in flight what limits throughput is how much of the GPU the latency-bound sampler
        # holds while it runs, not how long one call takes: the bucketed sampler (a quarter of an SM per scene
        # instead of four SMs) wherever it applies, fewer, longer-lived CTAs for the small fused-SA layers
        # (every CTA pays a fixed weight-staging cost) and one CTA per row tile for the 256-wide pm_linear layers
        # (tools/marginal_cost.py: the 14 pm_linear launches cost 77 us of a 412 us step, mostly CTA life time).
        from . import _ext
        if fps_algo is None:
            n = example.shape[1]
            fps_algo = _ext.FPS_BUCKET if _ext.FPS_BUCKET_MIN_N <= n <= _ext.FPS_BUCKET_MAX_N else _ext.FPS_AUTO
        self._options = dict(sa_min_tiles=int(sa_min_tiles), fps_algo=int(fps_algo), pm_n_tile=int(pm_n_tile),
                             pm_tiles_per_cta=int(pm_tiles_per_cta))
        self.model = model
        self.device = example.device
        self.n = int(n_streams)
        self.result_keys = tuple(result_keys) if result_keys else None
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.n)]
        self.static_in = [torch.empty_like(example) for _ in range(self.n)]
        self.graphs, self.outputs, self.host_out, self.done = [], [], [], []
        self.packed, self.host_packed = [], []
        self._next = 0
        self._pending_host = [False] * self.n
        self._submitted = [False] * self.n
        for s, x in zip(self.streams, self.static_in):
            x.copy_(example)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s), torch.no_grad(), _ext.launch_options(**self._options):
                for _ in range(warmup):                      # lazy inits (weight folding, func attrs)
                    model({"point_clouds": x})
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s), torch.no_grad(), _ext.launch_options(**self._options):
                out = model({"point_clouds": x})
                # the requested results are packed into ONE byte buffer inside the graph, so that a step
                # needs a single device->host copy instead of one per tensor (8 small copies per step
                # cost ~8 % of the pipeline's throughput on B200)
                packed = self._pack(out) if self.result_keys else None
            self.graphs.append(g)
            self.outputs.append(out)
            self.packed.append(packed)
            if self.result_keys:
                host = torch.empty(packed.shape, dtype=torch.uint8).pin_memory()
                self.host_packed.append(host)
                self.host_out.append(self._unpack(host, out))
            self.done.append(torch.cuda.Event())
        torch.cuda.synchronize(self.device)

    _ALIGN = 16

    def _segments(self, out):
        off, segs = 0, []
        for k in self.result_keys:
            nbytes = out[k].numel() * out[k].element_size()
            segs.append((k, off, nbytes))
            off += (nbytes + self._ALIGN - 1) // self._ALIGN * self._ALIGN
        return segs, off

    def _pack(self, out):
        segs, total = self._segments(out)
        buf = torch.empty(total, dtype=torch.uint8, device=self.device)
        for k, off, nbytes in segs:
            buf[off:off + nbytes].copy_(out[k].contiguous().view(-1).view(torch.uint8))
        return buf

    def _unpack(self, host, out):
        segs, _ = self._segments(out)
        return {k: host[off:off + nbytes].view(out[k].dtype).view(out[k].shape) for k, off, nbytes in segs}

    def submit(self, x, to_host=False):
        """Queue one batch on the next slot.  A device-resident `x` may still be being written on the caller's
        current stream (e.g. by DeviceSceneStore.make_batch): the slot stream waits for that stream first.
        With to_host=True the slot's pinned result buffer is overwritten, so the previous submission of the slot
        must have been consumed with wait() (checked)."""
        i = self._next
        if to_host and self._pending_host[i]:
            raise RuntimeError("GraphedDetector.submit: slot %d still holds unread host results; call wait(%d) first" % (i, i))
        self._next = (i + 1) % self.n
        s = self.streams[i]
        if x.is_cuda:
            s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self.static_in[i].copy_(x, non_blocking=True)
            self.graphs[i].replay()
            if to_host and self.result_keys:
                self.host_packed[i].copy_(self.packed[i], non_blocking=True)
                self._pending_host[i] = True
            self.done[i].record(s)
        self._submitted[i] = True
        return i

    def busy(self, i):
        """True when slot i has a submission whose host results have not been collected with wait()."""
        return self._pending_host[i]

    def wait(self, i):
        self.done[i].synchronize()
        self._pending_host[i] = False
        return self.host_out[i] if self.host_out else self.outputs[i]

    def close(self):
        """Drain every stream."""
        self.wait_all()
        self._pending_host = [False] * self.n

    def wait_all(self):
        for s in self.streams:
            s.synchronize()

    def fork_from(self, event):
        """Make every stream wait for `event` (start of a timed region)."""
        for s in self.streams:
            s.wait_event(event)

    def join_into(self, stream):
        """Make `stream` wait for everything submitted so far (end of a timed region)."""
        for s in self.streams:
            e = torch.cuda.Event()
            e.record(s)
            stream.wait_event(e)
