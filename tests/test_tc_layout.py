"""Index arithmetic of the chunked backward's shared-memory tiles (rwkvtts_b200/csrc/wkv7_tc_bwd.cu), restated on the CPU:
the stage-A thread mapping is a bijection onto (token, channel quad) and every access pattern of the stage is free of bank
conflicts; the padded 16 x 16 tile strides give the conflict degrees DESIGN.md 4.2 states; "G at the chunk start" lives in
gaps of the Gt tile that no tile element occupies.  (The constants are read out of the .cu file, so a changed stride or
mapping has to keep these properties.)"""
import os
import re
from collections import Counter

SRC = os.path.join(os.path.dirname(__file__), "..", "rwkvtts_b200", "csrc", "wkv7_tc_bwd.cu")


def consts():
    text = open(SRC).read()
    out = {}
    for name in ("N32_LBO", "N16_LBO", "N_SBO", "T_SBO", "T_LBO", "G_LBO", "S32_LBO", "S32T_LBO", "S16_LBO", "S_SBO"):
        m = re.search(r"\b%s = (\d+)" % name, text)
        assert m, name
        out[name] = int(m.group(1))
    assert "const int t = 4 * (wp >> 1) + tt, k4 = (lane & 7) + 8 * ((tt ^ wp) & 1);" in text       # the mapping restated below
    assert "return (row >> 3) * T_SBO + ((row & 7) >> 1) * T_LBO + 32 + (row & 1);" in text          # gs_off restated below
    return out


def kmajor(r, k, lbo, sbo):
    return (r >> 3) * sbo + (k >> 2) * lbo + (r & 7) * 4 + (k & 3)


def ways(word_addrs):
    """Worst number of distinct 32-bit words of one access that fall into the same bank."""
    c = Counter(a % 32 for a in set(word_addrs))
    return max(c.values())


def stage_a_lanes(wp):
    return [(4 * (wp >> 1) + (lane >> 3), (lane & 7) + 8 * (((lane >> 3) ^ wp) & 1)) for lane in range(32)]


def test_stage_a_mapping_is_a_bijection_and_conflict_free():
    K = consts()
    seen = set()
    for wp in range(8):
        lanes = stage_a_lanes(wp)
        seen.update(lanes)
        for j in range(4):           # scalar stores into the transposed [channel][token] tiles (Q~ A~ B~ K~ / dY G)
            for lbo in (K["G_LBO"], K["T_LBO"]):
                assert ways([(k4 >> 1) * K["T_SBO"] + (t >> 2) * lbo + (k4 & 1) * 16 + (t & 3) + 4 * j for t, k4 in lanes]) == 1
        for half in range(2):        # 8-byte reads of the landed [token][64] bf16 tiles, one half-warp per wavefront
            words = []
            for t, k4 in lanes[16 * half:16 * half + 16]:
                w = (t * 128 + k4 * 8) // 4
                words += [w, w + 1]
            assert ways(words) == 1
        for q in range(4):           # 16-byte stores into the [token][channel] tiles and the scan scratch, per quarter-warp
            for lbo in (K["N32_LBO"], K["N16_LBO"]):
                words = []
                for t, k4 in lanes[8 * q:8 * q + 8]:
                    o = (t >> 3) * K["N_SBO"] + k4 * lbo + (t & 7) * 4
                    words += [o, o + 1, o + 2, o + 3]
                assert ways(words) == 1
            words = []
            for t, k4 in lanes[8 * q:8 * q + 8]:
                words += [t * 64 + k4 * 4 + e for e in range(4)]
            assert ways(words) == 1
    assert seen == {(t, k4) for t in range(16) for k4 in range(16)}


def test_padded_tile_strides_give_the_documented_conflict_degrees():
    K = consts()
    perm8 = lambda g: (g & 1) | ((g & 2) << 1) | ((g & 4) >> 1)
    gram, col = [], []
    for nt in range(2):
        for e in range(2):
            for hh in range(2):      # stage B: fragment stores of a Gram block into a [s][t] operand tile
                gram.append(ways([kmajor(2 * perm8(2 * (lane & 3) + e) + nt, 2 * perm8(lane >> 2) + hh, K["S16_LBO"], K["S_SBO"])
                                  for lane in range(32)]))
    for s in range(16):              # stage B: one column per thread
        col.append(ways([kmajor(s, t, K["S16_LBO"], K["S_SBO"]) for t in range(16)]))
    assert max(gram) == 2 and max(col) == 1
    nat, trn = [], []
    for q in range(4):               # group C1: gradient Gram blocks, natural and transposed copies
        for nt in range(2):
            for e in range(2):
                for hh in range(2):
                    a_n, a_t = [], []
                    for lane in range(32):
                        g, tq = lane >> 2, lane & 3
                        c, r = 8 * nt + 2 * tq + e, g + 8 * hh
                        a_n.append(kmajor((q & 1) * 16 + r, c, K["S32_LBO"], K["S_SBO"]))
                        a_t.append(kmajor((q >> 1) * 16 + c, r, K["S32T_LBO"], K["S_SBO"]))
                    nat.append(ways(a_n))
                    trn.append(ways(a_t))
    assert max(trn) == 1 and max(nat) == 2


def test_g_at_chunk_start_sits_in_gaps_of_the_gt_tile():
    K = consts()
    gs_off = lambda row: (row >> 3) * K["T_SBO"] + ((row & 7) >> 1) * K["T_LBO"] + 32 + (row & 1)
    tile = {kmajor(ch, tok, K["T_LBO"], K["T_SBO"]) for ch in range(64) for tok in range(16)}
    slots = [gs_off(r) for r in range(64)]
    assert len(set(slots)) == 64 and not (set(slots) & tile) and max(slots) < 4 * K["T_LBO"]
