"""Host logic of the beam-search launch plan (gbnns_dim_red_b200/csrc/beam_search.cu: beam_plan), checked on a grid of
(ef, dimension, index size) without a GPU: the plan must fit the SM it is made for and keep the exactness conditions of
the visited-set formats."""
import numpy as np
import pytest

from gbnns_dim_red_b200 import capi

SM_SHARED = 228 * 1024        # bytes of shared memory per SM (1 KB of it reserved per CTA, counted in smem_per_sm)
CTA_SHARED = 227 * 1024       # opt-in maximum per CTA
# resident warps the kernels' register budgets allow (16-bit-tag builds; v2_regs / v2_shape in csrc/beam_search.cuh):
# a whole number of warps per scheduler partition of the register file
REG_WARPS = {32: 32, 64: 32, 96: 28, 128: 24, 160: 24, 192: 20, 256: 16, 320: 16, 384: 16, 512: 12}
REG_WARPS_32BIT = {32: 24, 64: 24, 96: 20, 128: 20, 160: 20, 192: 16, 256: 16, 320: 12, 384: 12, 512: 12}

EFS = [1, 2, 8, 24, 25, 53, 56, 57, 87, 88, 100, 120, 121, 174, 175, 248, 249, 294, 330, 400, 500, 504]


@pytest.mark.parametrize("n", [64, 3000, 100_000, 1_000_000, 4_000_000, 12_500_000, 100_000_000])
@pytest.mark.parametrize("dim", [16, 32, 48, 64])
def test_plans_fit_the_sm_and_stay_exact(n, dim):
    for ef in EFS:
        p = capi.beam_plan_info(ef, dim, n)
        assert p["variant"] == 2, "d_low in {16,32,48,64} and ef <= 504 run in the batched-merge kernel"
        assert p["cap"] in REG_WARPS and p["cap"] >= ef + 8, (ef, p)
        warps = p["warps_per_cta"] * p["ctas_per_sm"]
        assert 1 <= p["ctas_per_sm"] <= 32 and 1 <= p["warps_per_cta"] <= 32
        assert p["cap"] == min(c for c in REG_WARPS if c >= ef + 8), "smallest capacity that holds ef + tie slack"
        assert p["smem_per_sm"] <= SM_SHARED, (n, dim, ef, p)
        assert p["smem_per_warp"] * p["warps_per_cta"] <= CTA_SHARED
        assert p["smem_per_warp"] % 16 == 0 and p["vis_bytes"] % 16 == 0
        if p["tag_bits"]:
            # 16-bit tags: (bucket, tag) must identify the id, tag + displacement bits fit 15 bits, and the register
            # budget of the list capacity bounds the resident warps
            buckets = p["vis_bytes"] // 16
            b = max(1, int(np.ceil(np.log2(n))))
            assert p["vis_entries"] == 7 * buckets
            assert p["tag_bits"] == b - int(np.floor(np.log2(buckets))) <= 14, (n, ef, p)
            assert 1 <= p["disp_bits"] <= 2 and p["tag_bits"] + p["disp_bits"] <= 15
            dense = p["cap"] <= 64 and (p["warps_per_cta"], p["ctas_per_sm"]) == (17, 2)   # the 56-register build
            assert dense or warps <= REG_WARPS[p["cap"]], (ef, p)
            assert dense or p["ctas_per_sm"] == 1 or p["warps_per_cta"] % 4 == 0, "CTAs spread evenly over the four schedulers"
            assert 4 * (12 * ef + 200) <= 3 * p["vis_entries"], "expected visited count above 75 % of the table"
        else:
            assert p["vis_entries"] * 4 == p["vis_bytes"]
            assert warps <= REG_WARPS_32BIT[p["cap"]], (ef, p)
        assert p["vis_entries"] >= 64


def test_headline_shape_plan():
    """SIFT-1M at the bench's operating point: 2 CTAs x 17 warps of the 56-register build (34 resident warps: a 10 000-query
    batch is two full waves on 148 SMs) with a 233-bucket visited table each."""
    p = capi.beam_plan_info(53, 32, 1_000_000)
    assert (p["cap"], p["warps_per_cta"], p["ctas_per_sm"]) == (64, 17, 2)
    assert p["smem_per_sm"] <= SM_SHARED and p["vis_bytes"] == 16 * 233 and p["tag_bits"] == 13
    # the larger lists keep the tag format and trade warps for table size
    assert capi.beam_plan_info(100, 32, 1_000_000)["warps_per_cta"] * capi.beam_plan_info(100, 32, 1_000_000)["ctas_per_sm"] == 24
    p200 = capi.beam_plan_info(200, 32, 1_000_000)
    assert p200["ctas_per_sm"] * p200["warps_per_cta"] == 16
    # no cliff between list capacities: 140 gets a 160-slot list at the residency of the 128-slot one
    p140 = capi.beam_plan_info(140, 32, 1_000_000)
    assert p140["cap"] == 160 and p140["warps_per_cta"] * p140["ctas_per_sm"] >= 20
    # the two-graph mode runs in the shared-memory-list kernel
    assert capi.beam_plan_info(53, 32, 1_000_000, second_graph=True)["variant"] == 0
    # dimensions the batched-merge kernel does not cover fall back to the sequential register kernel
    assert capi.beam_plan_info(53, 24, 1_000_000)["variant"] == 1


def test_large_shards_keep_the_tag_format_where_it_pays():
    """A 12.5 M-vertex shard (Deep-100M over 8 GPUs) needs >= 1024 buckets for 14-bit tags: taken for the wide beams whose
    table is that large anyway, while small beams keep more warps resident with 32-bit slots."""
    wide = capi.beam_plan_info(376, 16, 12_500_000)
    assert wide["tag_bits"] == 14 and wide["disp_bits"] == 1 and wide["vis_bytes"] >= 16 * 1024
    small = capi.beam_plan_info(53, 16, 12_500_000)
    assert small["tag_bits"] == 0 and small["warps_per_cta"] * small["ctas_per_sm"] >= 20
