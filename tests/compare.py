"""Parity metrics between the CUDA renderer and the oracle (BASELINE.json north_star):

  * DEM decode/mosaic: bit-exact (checked elsewhere)
  * pixel coverage (terrain vs sky) must agree on >= 99.5 % of the pixels
  * where both see terrain, range must agree within 1e-4 relative; pixels that do not are only
    tolerated on silhouette edges (a depth discontinuity or a terrain/sky boundary within one
    pixel in the oracle image) and count against the same 0.5 % budget
  * red channel within +-1 LSB wherever range agrees
"""
import numpy as np

RANGE_RTOL = 1e-4          # north_star: "range must agree within 1e-4 relative where both renderers hit terrain"
MIN_AGREEMENT = 0.995      # north_star: ">= 99.5 % of pixels"
SILHOUETTE_JUMP = 1e-3     # relative range step between neighbours that marks a depth discontinuity


def silhouette_mask(ranges):
    """Pixels within one pixel of a terrain/sky boundary or of a relative range jump > SILHOUETTE_JUMP."""
    r = ranges.astype(np.float64)
    H, W = r.shape
    pad = np.pad(r, 1, mode="edge")
    mask = np.zeros((H, W), bool)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            nb = pad[1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
            hit, nbhit = r > 0, nb > 0
            mask |= hit != nbhit
            both = hit & nbhit
            jump = np.zeros_like(mask)
            jump[both] = np.abs(nb[both] - r[both]) > SILHOUETTE_JUMP * np.minimum(nb[both], r[both])
            mask |= jump
    return mask


def compare_renders(img, rng, img_ref, rng_ref):
    """Returns a dict of parity statistics; `ok` says whether the north_star bar is met."""
    assert img.shape == img_ref.shape and rng.shape == rng_ref.shape
    n = rng.size
    hit, hit_ref = rng > 0, rng_ref > 0
    coverage_agree = float((hit == hit_ref).sum()) / n
    both = hit & hit_ref
    rel = np.zeros(rng.shape)
    rel[both] = np.abs(rng[both].astype(np.float64) - rng_ref[both]) / rng_ref[both]
    range_bad = both & (rel > RANGE_RTOL)
    sil = silhouette_mask(rng_ref)
    bad = (hit != hit_ref) | range_bad
    bad_off_silhouette = bad & ~sil
    red_diff = np.abs(img[..., 2].astype(int) - img_ref[..., 2].astype(int))
    good = both & ~range_bad
    sky = ~hit & ~hit_ref
    stats = dict(
        pixels=n,
        hit_fraction_ref=float(hit_ref.mean()),
        coverage_agreement=coverage_agree,
        range_mismatch=int(range_bad.sum()),
        agreement=1.0 - float(bad.sum()) / n,
        off_silhouette=int(bad_off_silhouette.sum()),
        max_rel_range_err_where_agree=float(rel[good].max()) if good.any() else 0.0,
        red_max_diff_where_agree=int(red_diff[good].max()) if good.any() else 0,
        bit_exact=bool((img == img_ref).all() and (rng == rng_ref).all()),
        sky_ok=bool((img[sky] == (255, 0, 0)).all() and (rng[sky] == -1.0).all()),
        bg_channels_ok=bool((img[..., 1] == 0).all()),
    )
    stats["ok"] = (stats["coverage_agreement"] >= MIN_AGREEMENT and stats["agreement"] >= MIN_AGREEMENT
                   and stats["off_silhouette"] <= max(2, n // 20000) and stats["red_max_diff_where_agree"] <= 1
                   and stats["sky_ok"] and stats["bg_channels_ok"])
    return stats
