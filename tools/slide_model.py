"""Symbolic model of the sliding-window filter stage (stage D of the pipelined kernel): checks, without any floating point,
that every lane multiplies exactly the (tap, coefficient) pairs of the reference's chain order (DotProdPatch_AVX512_32f:
chain j accumulates taps 16m + j, m = 0..7, then the 16-lane tree) and that the folded lane tree adds the same operands in
the same shape.  Also prints the packed shuffle-source table the kernel uses (c_slide_tbl).
usage: python tools/slide_model.py"""
import itertools


def t_of(N):            # lanes q < t_of(N) are one pair behind ("flagged"); chain pair of lane q is (q - t) & 7
    return (3 * N) % 8


def base_of(N):         # first pair (of the lane's own pair list) of the uniform window at step N
    return (11 * N) // 8


def check_macs(nsteps=24):
    for N in range(nsteps):
        t, B = t_of(N), base_of(N)
        for q in range(8):
            flag = q < t
            p = (q - t) & 7
            for m in range(8):
                i = B + (1 if flag else 0) + m            # pair of the lane's list used with coefficient unit m
                for e in range(2):
                    k0 = 16 * i + 2 * q + e               # position in the linearised column strip (11 R + c)
                    kn = k0 - 22 * N                      # tap of pixel N (two rows = 22 taps further down per step)
                    assert kn == 16 * m + 2 * p + e, (N, q, m, e, kn)
            # the uniform window [B, B + 8] covers both cases; at t == 0 only [B, B + 7] is needed
    # ring of 11: pairs loaded for step N + 1 never overwrite a pair step N still reads
    for N in range(nsteps):
        L = lambda n: 7 + 11 * (n // 8) if n % 8 == 0 else base_of(n) + 8
        new = range(L(N) + 1, L(N + 1) + 1)
        live = range(base_of(N), base_of(N) + (8 if N % 8 == 0 else 9))
        assert not ({i % 11 for i in new} & {i % 11 for i in live}), N
    print("MAC schedule ok")


def tree_tables():
    words = []
    for q in range(8):
        w = 0
        for U, n0 in enumerate((0, 4)):
            t = [t_of(n0 + u) for u in range(4)]
            pi = lambda u, lane: ((((lane - t[u]) & 7) ^ 2) + t[u]) & 7
            s2a = pi(2, q) if q & 2 else pi(0, q)
            s2b = pi(3, q) if q & 2 else pi(1, q)
            u = q & 3
            cands = [q ^ 1, q ^ 5]
            s3 = [r for r in cands if (((r - t[u]) ^ (q - t[u])) & 4) == 0]
            assert len(s3) == 1
            s3 = s3[0]
            hi = sum((1 if ((q - t[k]) & 4) else 0) << k for k in range(4))
            field = s2a | (s2b << 3) | (s3 << 6) | (hi << 9)      # 13 bits per unit type
            w |= field << (16 * U)
        words.append(w)
    return words


def check_tree(words):
    # symbolic: a value is a nested frozenset/tuple structure; fadd(a, b) -> ('+', frozenset({a, b})) (commutative)
    add = lambda a, b: ('+', frozenset((a, b)))
    for U, n0 in enumerate((0, 4)):
        # per pixel u, per lane q: chain sums (a0, a1) of chain pair p
        a = [[None] * 8 for _ in range(4)]
        for u in range(4):
            t = t_of(n0 + u)
            for q in range(8):
                p = (q - t) & 7
                a[u][q] = (('c', u, 2 * p), ('c', u, 2 * p + 1))
        f = lambda q, sh, n: (words[q] >> (16 * U + sh)) & ((1 << n) - 1)
        v = [[None] * 8 for _ in range(4)]
        for u in range(4):
            for q in range(8):
                hi = (f(q, 9, 4) >> u) & 1
                hi_src = (f(q ^ 4, 9, 4) >> u) & 1
                sent = a[u][q ^ 4][0] if hi_src else a[u][q ^ 4][1]
                v[u][q] = add(a[u][q][1] if hi else a[u][q][0], sent)
        w0, w1 = [None] * 8, [None] * 8
        for q in range(8):
            b1 = (q >> 1) & 1
            s = f(q, 0, 3); sb1 = (s >> 1) & 1
            w0[q] = add(v[2][q] if b1 else v[0][q], v[0][s] if sb1 else v[2][s])
            s = f(q, 3, 3); sb1 = (s >> 1) & 1
            w1[q] = add(v[3][q] if b1 else v[1][q], v[1][s] if sb1 else v[3][s])
        x = [None] * 8
        for q in range(8):
            b0 = q & 1
            s = f(q, 6, 3); sb0 = s & 1
            x[q] = add(w1[q] if b0 else w0[q], w0[s] if sb0 else w1[s])
        cur = [add(x[q], x[q ^ 4]) for q in range(8)]
        for q in range(8):
            u = q & 3
            c = lambda j: ('c', u, j)
            t8 = [add(c(j), c(j + 8)) for j in range(8)]
            t4 = [add(t8[j], t8[j + 4]) for j in range(4)]
            t2 = [add(t4[j], t4[j + 2]) for j in range(2)]
            assert cur[q] == add(t2[0], t2[1]), (U, q)
    print("folded tree ok")


if __name__ == "__main__":
    check_macs()
    words = tree_tables()
    check_tree(words)
    print("c_slide_tbl = {" + ", ".join("0x%08xu" % w for w in words) + "}")
