"""Second, independent oracle: the declarative spec of SURVEY.md section 8c, by brute force.

Occ = {(pid, s, e) : haystack[s:e] == pattern[pid]} (ASCII case folded if enabled), restricted to
the span.  O(n * sum|p|): for small test inputs only.
"""
STANDARD, LEFTMOST_FIRST, LEFTMOST_LONGEST = 0, 1, 2


def _fold(b: bytes) -> bytes:
    return bytes(c + 32 if 65 <= c <= 90 else c for c in b)


def occurrences(patterns, hay, span=None, ci=False):
    s0, e0 = span if span is not None else (0, len(hay))
    h = _fold(hay) if ci else bytes(hay)
    occ = []
    for pid, p in enumerate(patterns):
        p = _fold(p) if ci else bytes(p)
        for s in range(s0, e0 - len(p) + 1):
            if h[s:s + len(p)] == p:
                occ.append((pid, s, s + len(p)))
    return occ


def find_overlapping(patterns, hay, span=None, ci=False):
    occ = occurrences(patterns, hay, span, ci)
    return sorted(occ, key=lambda m: (m[2], -(m[2] - m[1]), m[0]))


def find_iter(patterns, hay, kind, span=None, ci=False):
    s0, e0 = span if span is not None else (0, len(hay))
    occ = occurrences(patterns, hay, span, ci)
    if kind == STANDARD:
        key = lambda m: (m[2], -(m[2] - m[1]), m[0])
    elif kind == LEFTMOST_FIRST:
        key = lambda m: (m[1], m[0])
    else:
        key = lambda m: (m[1], -(m[2] - m[1]), m[0])
    out, p, last_end = [], s0, None
    while p <= e0:
        cand = [m for m in occ if m[1] >= p]
        if not cand:
            break
        m = min(cand, key=key)
        if m[1] == m[2] and last_end is not None and m[2] == last_end:
            p += 1
            cand = [x for x in occ if x[1] >= p]
            if not cand:
                break
            m = min(cand, key=key)
        out.append(m)
        p = m[2]
        last_end = m[2]
    return out
