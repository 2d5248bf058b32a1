"""Work distribution of the tensor-core screen (csrc/topk_screen.cuh: ScreenParams, screen_item_at; topk.cu:
topk_rerank_kernel's item discovery), restated in Python and checked exhaustively on small shapes: every (query tile,
dictionary tile) unit is covered exactly once, the CTAs' loads differ by at most one unit, item ids `cta + qt` are
unique and below n_ctas + n_qtiles, and the re-rank finds exactly the items the screen wrote."""
import itertools


def span_begin(c, n_qtiles, T, n_ctas):
    """screen_span_begin: the first U % n_ctas CTAs take one unit more than U // n_ctas."""
    units = n_qtiles * T
    base, rem = units // n_ctas, units % n_ctas
    return c * base + min(c, rem)


def span_owner(u, n_qtiles, T, n_ctas):
    """screen_span_owner: the CTA whose span holds unit u."""
    units = n_qtiles * T
    base, rem = units // n_ctas, units % n_ctas
    big = rem * (base + 1)
    if u < big:
        return u // (base + 1)
    return rem + (u - big) // max(base, 1)


def screen_items(n_qtiles, T, n_ctas):
    """What topk_screen_kernel's roles enumerate: per CTA the list of (id, qt, t0, t1)."""
    out = []
    for c in range(n_ctas):
        u, u_end = span_begin(c, n_qtiles, T, n_ctas), span_begin(c + 1, n_qtiles, T, n_ctas)
        items = []
        while u < u_end:
            qt = u // T
            t0 = u - qt * T
            t1 = min(T, t0 + (u_end - u))
            items.append((c + qt, qt, t0, t1))
            u += t1 - t0
        out.append(items)
    return out


def rerank_items(qt, n_qtiles, T, n_ctas):
    """What topk_rerank_kernel enumerates for a query of tile qt: (id, t0, t1)."""
    u_lo, u_hi = qt * T, (qt + 1) * T
    c = span_owner(u_lo, n_qtiles, T, n_ctas)
    assert span_begin(c, n_qtiles, T, n_ctas) <= u_lo < span_begin(c + 1, n_qtiles, T, n_ctas)
    found = []
    while c < n_ctas:
        s0, s1 = span_begin(c, n_qtiles, T, n_ctas), span_begin(c + 1, n_qtiles, T, n_ctas)
        if s0 >= u_hi:
            break
        lo, hi = max(s0, u_lo), min(s1, u_hi)
        if lo < hi:
            found.append((c + qt, lo - u_lo, hi - u_lo))
        c += 1
    return found


def test_balanced_spans_cover_every_unit_once_and_rerank_finds_them():
    shapes = list(itertools.product([1, 2, 3, 7, 79, 625], [1, 2, 5, 8, 33, 611], [1, 3, 8, 148]))
    for n_qtiles, T, sms in shapes:
        units = n_qtiles * T
        n_ctas = sms if units >= sms else max(1, units)
        per_cta = screen_items(n_qtiles, T, n_ctas)
        covered = {}
        ids = {}
        for c, items in enumerate(per_cta):
            load = sum(t1 - t0 for _, _, t0, t1 in items)
            assert units // n_ctas <= load <= units // n_ctas + 1
            for iid, qt, t0, t1 in items:
                assert 0 <= iid < sms + n_qtiles and t0 < t1 <= T
                assert iid not in ids, "item ids must be unique"
                ids[iid] = (qt, t0, t1)
                for t in range(t0, t1):
                    assert (qt, t) not in covered
                    covered[(qt, t)] = iid
        assert len(covered) == units
        for qt in range(n_qtiles):
            want = sorted((iid, t0, t1) for iid, (q, t0, t1) in ids.items() if q == qt)
            assert sorted(rerank_items(qt, n_qtiles, T, n_ctas)) == want
