"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Checkers for the code-table builder. The reference has no table builder (it only consumes .def
tables), so there is nothing of the reference to restate or pin to: "parity unpinned" by necessity. The checks
are the published algorithms written a second time, independently of the product's C:
  * huffman_cost        classic Huffman (heap): the optimal cost without a length limit
  * package_merge       Larmore-Hirschberg package-merge on (weight, symbols) tuples: optimal lengths under a limit
  * brute_force_cost    every length assignment that satisfies Kraft, for tiny alphabets (pins the two above)
  * canonical           canonical code assignment by (length, symbol)
"""
import heapq
import itertools
from fractions import Fraction


def huffman_cost(weights):
    w = [x for x in weights if x > 0]
    if len(w) <= 1:
        return sum(w)
    heapq.heapify(w)
    cost = 0
    while len(w) > 1:
        a, b = heapq.heappop(w), heapq.heappop(w)
        cost += a + b
        heapq.heappush(w, a + b)
    return cost


def package_merge(weights, max_bits):
    """weights: list of (weight, symbol), all taking part. Returns {symbol: length}."""
    n = len(weights)
    if n == 0:
        return {}
    if n == 1:
        return {weights[0][1]: 1}
    assert (1 << max_bits) >= n
    leaves = sorted(((w, (s,)) for w, s in weights), key=lambda t: (t[0], t[1]))
    cur = list(leaves)
    for _ in range(max_bits - 1):
        packages = [(cur[2 * i][0] + cur[2 * i + 1][0], cur[2 * i][1] + cur[2 * i + 1][1]) for i in range(len(cur) // 2)]
        merged, li, pi = [], 0, 0
        while li < len(leaves) or pi < len(packages):
            if pi >= len(packages) or (li < len(leaves) and leaves[li][0] <= packages[pi][0]):
                merged.append(leaves[li])
                li += 1
            else:
                merged.append(packages[pi])
                pi += 1
        cur = merged
    lengths = {s: 0 for _, s in weights}
    for _, syms in cur[:2 * n - 2]:
        for s in syms:
            lengths[s] += 1
    return lengths


def brute_force_cost(weights, max_bits):
    """Minimum of sum(w * len) over all length vectors with Kraft sum <= 1 (tiny alphabets only)."""
    n = len(weights)
    if n == 1:
        return weights[0]
    best = None
    for lens in itertools.product(range(1, max_bits + 1), repeat=n):
        if sum(Fraction(1, 1 << l) for l in lens) <= 1:
            c = sum(w * l for w, l in zip(weights, lens))
            best = c if best is None else min(best, c)
    return best


def canonical(lengths, eos_length=0):
    """lengths: 256 ints. Returns ({symbol: (pattern, len)}, eos (pattern, len) or None)."""
    order = sorted((l, s) for s, l in enumerate(list(lengths) + [eos_length]) if l)
    codes, code, prev = {}, 0, 0
    for l, s in order:
        code <<= l - prev
        prev = l
        codes[s] = (code, l)
        code += 1
    eos = codes.pop(256, None)
    return codes, eos
