"""Generates (and checks) the comparator list of `median27` in v-express_b200/csrc/vx_post.cu:
Batcher odd-even merge sort for 32 wires, minus comparators touching the 5 (+inf) padding wires, pruned backwards to the
comparators wire 13 (the 14th smallest of 27) depends on.   usage: python profiles/tools/gen_median_network.py"""
import random


def batcher(n):
    pairs, p = [], 1
    while p < n:
        k = p
        while k >= 1:
            j = k % p
            while j + k < n:
                for i in range(k):
                    if i + j + k < n and (i + j) // (2 * p) == (i + j + k) // (2 * p):
                        pairs.append((i + j, i + j + k))
                j += 2 * k
            k //= 2
        p *= 2
    return pairs


pairs = [(a, b) for a, b in batcher(32) if b < 27]
need, keep = {13}, []
for a, b in reversed(pairs):
    if a in need or b in need:
        keep.append((a, b))
        need |= {a, b}
keep.reverse()
for _ in range(20000):
    v = [random.random() for _ in range(27)]
    if random.random() < 0.3:
        v = [random.choice([0.0, 0.25, 0.5, 1.0]) for _ in range(27)]
    w = list(v)
    for a, b in keep:
        if w[a] > w[b]:
            w[a], w[b] = w[b], w[a]
    assert w[13] == sorted(v)[13]
print(f"// {len(keep)} comparators")
for i in range(0, len(keep), 10):
    print("  " + " ".join(f"CS({a},{b})" for a, b in keep[i:i + 10]))
