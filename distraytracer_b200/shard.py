"""Work partitioning across GPUs: samples are independent given an immutable,
replicated scene, so there is no exchange step -- frames (video) or tiles (single frame) are
dealt out and only the finished u8 pixels are gathered.  torch.distributed is used for the
barrier and the max-over-ranks timing, never on the data path."""
from . import abi


def frames_for_rank(rank, world, n_frames, start=0):
    """Video: frame f goes to rank f mod world (round-robin keeps neighbouring, similarly
    expensive frames on different GPUs)."""
    return [start + f for f in range(n_frames) if f % world == rank]


def bands_for_devices(yres, n):
    """Single frame: one horizontal band of loop rows per device, [(y0, y1), ...]; band d is
    rows [yres*d/n, yres*(d+1)/n) exactly as distraytracer_b200/host/drt_host.h cuts them."""
    return [(yres * d // n, yres * (d + 1) // n) for d in range(n)]


def interleaved_tiles(xres, yres, n, tile=64):
    """Single frame with strongly non-uniform cost (cloud sky vs. geometry): tile x tile blocks
    dealt round-robin, [[abi.Tile, ...] per device]."""
    out = [[] for _ in range(n)]
    k = 0
    for y0 in range(0, yres, tile):
        for x0 in range(0, xres, tile):
            out[k % n].append(abi.Tile(x0, y0, min(tile, xres - x0), min(tile, yres - y0), k % n))
            k += 1
    return out


def place_band(frame_rgb, band_rgb, y0, y1):
    """Gather: copy a band rendered for loop rows [y0,y1) into the PPM-ordered full frame."""
    yres = frame_rgb.shape[0]
    frame_rgb[yres - y1: yres - y0] = band_rgb


def row_blocks(yres, rows):
    """Single frame, dynamic partition: loop-row ranges [(y0, y1), ...] of at most `rows` rows, bottom to top."""
    rows = max(1, int(rows))
    return [(y0, min(y0 + rows, yres)) for y0 in range(0, yres, rows)]


def render_frame_blocks(handles, settings, frame_rgb, rows=15):
    """Render one frame on several GPUs: one host thread per handle claims row blocks from a shared counter and
    renders each block straight into its place in `frame_rgb` ((yRes, xRes, 3) uint8, PPM row order; pinned memory
    makes the device-to-host copies asynchronous).  `handles` are runtime.DeviceScene objects of the SAME scene --
    pass two per GPU: render_wave is a persistent kernel whose last batches drain with few warps busy, and a second
    stream lets the SMs a finished block frees start on the next block at once.  ctypes releases the GIL inside
    drt_render, so the threads overlap.  This is drt_host.h::renderFrame in Python.  Returns the wall-clock seconds."""
    import threading
    import time
    yres, xres = settings.yRes, settings.xRes
    from .runtime import check_frame_buffer
    check_frame_buffer(frame_rgb, yres, xres)
    blocks = row_blocks(yres, rows)
    nxt, lock, errs = [0], threading.Lock(), []

    def work(h):
        try:
            while True:
                with lock:
                    k = nxt[0]; nxt[0] += 1
                if k >= len(blocks):
                    return
                y0, y1 = blocks[k]
                h.render(settings, abi.Tile(0, y0, xres, y1 - y0, h.device), out=frame_rgb[yres - y1: yres - y0])
        except Exception as e:      # noqa: BLE001  (re-raised on the caller's thread)
            errs.append(e)

    threads = [threading.Thread(target=work, args=(h,)) for h in handles]
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errs:
        raise errs[0]
    return time.perf_counter() - t0


class FrameGroup:
    """One scene replicated on several GPUs of a box, for rendering single frames on all of them (BASELINE configs 3 and 5;
    SURVEY.md 8e).  `render` returns the wall-clock seconds of one frame, written into `frame_rgb` ((yRes, xRes, 3) uint8,
    PPM row order; pinned memory keeps the device-to-host copy asynchronous).

    method "steal": drt_render_multi -- one persistent kernel per GPU, all claiming ~1024-sample units from one counter
    in GPU 0's memory, pixels resolved into GPU 0's frame over NVLink, one copy to the host.
    method "blocks": the fallback without peer access -- row blocks claimed by host threads (render_frame_blocks)."""

    def __init__(self, scene, devices, method="steal", streams_per_gpu=2, rows=15):
        from .runtime import DeviceScene
        self.devices = list(devices)
        self.rows = rows
        self.kind = method
        per = 1 if method == "steal" else streams_per_gpu
        self.handles = [[DeviceScene(scene, d) for _ in range(per)] for d in self.devices]
        self.method = ("drt_render_multi: one launch per GPU, pixel-aligned ~1024-sample units stolen from one counter in GPU 0's "
                       "memory, resolved into GPU 0's frame over NVLink" if method == "steal" else
                       f"row blocks of {rows} rows claimed dynamically by one host thread per scene handle, "
                       f"{streams_per_gpu} handles (streams) per GPU")

    def render(self, settings, frame_rgb, gpus=None):
        import time
        use = self.handles[:gpus] if gpus else self.handles
        flat = [h for g in use for h in g]
        if self.kind == "steal":
            from .runtime import render_multi
            t0 = time.perf_counter()
            render_multi(flat, settings, out=frame_rgb)
            return time.perf_counter() - t0
        return render_frame_blocks(flat, settings, frame_rgb, self.rows)

    def close(self):
        for g in self.handles:
            for h in g:
                h.close()
        self.handles = []


def max_over_ranks(value, dist=None, device=None):
    """Timing contract of bench.py: the slowest rank defines the step time."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
