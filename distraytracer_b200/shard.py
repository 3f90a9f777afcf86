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


def max_over_ranks(value, dist=None, device=None):
    """Timing contract of bench.py: the slowest rank defines the step time."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
