"""Host-side sharding of the RAISR path across GPUs (no data-path collective in either scheme).

frame_shard   frame-parallel: frames are independent (RNLProcess is stateless per call, Raisr.cpp:1294-1397).
row_bands     row-band: contiguous bands of OUTPUT rows aligned to the upscale ratio -- the decomposition the reference
              applies across its threads (Raisr.cpp:1738-1779).  A band's result depends on output rows +-7 around it
              (6 for the patch/gradients, 1 for the census neighbours); instead of exchanging halo rows each rank reads
              the few extra INPUT rows it needs (band_input_rows) and recomputes them.
"""

HALO_OUT_ROWS = 7      # gLoopMargin (6) + CTmargin (1), Raisr_globals.h:33-36, Raisr.cpp:1574-1575


def frame_shard(rank, world, n_frames):
    """Frames rank `rank` of `world` processes: rank, rank+world, ... (round robin keeps a stream's latency even)."""
    return list(range(rank, n_frames, world))


def row_bands(out_h, world, align=2):
    """[(r0, r1)] per rank: contiguous, covering [0, out_h), starts aligned to `align` output rows."""
    base = (out_h // world) // align * align
    bands, r = [], 0
    for i in range(world):
        r1 = out_h if i == world - 1 else r + base
        bands.append((r, r1))
        r = r1
    return bands


def band_input_rows(r0, r1, ratio, in_h, out_h):
    """Input rows [a, b) that output rows [r0, r1) depend on (single-pass configurations): the +-7 output-row halo
    mapped through the pixel-centre upscale, plus one row each side for the bilinear taps."""
    lo = max(0, r0 - HALO_OUT_ROWS)
    hi = min(out_h, r1 + HALO_OUT_ROWS)
    a = int((lo + 0.5) / ratio - 0.5) - 1
    b = int((hi - 1 + 0.5) / ratio - 0.5) + 3
    return max(0, a), min(in_h, b)
