"""Block-by-block comparison of two .sla streams of the same file (host-side Python, no codec logic):
the report north_star asks for - "any block whose quantised coefficients differ must be listed and its
compressed-size delta reported".  The container is walked by its size fields; the oracle is
not needed."""
from __future__ import annotations

HEADER = 43


def walk(stream: bytes):
    """[(byte offset, block size, samples, block type)] by following the size fields"""
    out, off = [], HEADER
    total = int.from_bytes(stream[15:19], "big")
    done = 0
    while done < total and off + 11 <= len(stream):
        if stream[off] != 0xFF or stream[off + 1] != 0xFF:
            raise ValueError(f"no sync code at byte {off}")
        size = int.from_bytes(stream[off + 2:off + 6], "big") + 6
        n = int.from_bytes(stream[off + 8:off + 10], "big")
        out.append((off, size, n, stream[off + 10] >> 6))
        off += size
        done += n
    return out


def diff_streams(mine: bytes, ref: bytes) -> dict:
    """Aligns the two block chains on sample positions.  Returns
    {"identical": bool, "blocks": (mine, ref), "bytes": (mine, ref), "size_delta_ratio": float,
     "header_equal": bool, "mismatches": [{"sample_offset", "samples", "mine": {...}, "ref": {...}, "size_delta"}]}
    where a mismatch is a run of samples over which the two chains differ (block boundaries, types or bytes);
    a run ends where both chains reach the same sample position again."""
    a, b = walk(mine), walk(ref)
    res = {"identical": mine == ref, "blocks": (len(a), len(b)), "bytes": (len(mine), len(ref)),
           "size_delta_ratio": (len(mine) - len(ref)) / max(len(ref), 1),
           "header_equal": mine[:HEADER] == ref[:HEADER], "mismatches": []}
    i = j = 0
    pa = pb = 0            # sample positions
    while i < len(a) and j < len(b):
        oa, sa, na, ta = a[i]
        ob, sb, nb, tb = b[j]
        if pa == pb and na == nb and mine[oa:oa + sa] == ref[ob:ob + sb]:
            i += 1; j += 1; pa += na; pb += nb
            continue
        # a differing run: advance whichever chain is behind until they meet again
        start, i0, j0 = min(pa, pb), i, j
        ea, eb = pa + na, pb + nb
        i += 1; j += 1
        while ea != eb:
            if ea < eb and i < len(a):
                ea += a[i][2]; i += 1
            elif j < len(b):
                eb += b[j][2]; j += 1
            else:
                break
        pa, pb = ea, eb
        ma, mb = a[i0:i], b[j0:j]
        res["mismatches"].append({
            "sample_offset": start, "samples": max(ea, eb) - start,
            "mine": {"blocks": [(x[2], x[3], x[1]) for x in ma], "bytes": sum(x[1] for x in ma)},
            "ref": {"blocks": [(x[2], x[3], x[1]) for x in mb], "bytes": sum(x[1] for x in mb)},
            "size_delta": sum(x[1] for x in ma) - sum(x[1] for x in mb)})
    return res
