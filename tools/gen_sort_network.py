"""Batcher's odd-even merge sort network for 16 inputs, as used by match_tc.cu to find the k-th largest
partition maximum; verified on all 2^16 zero-one inputs (0-1 principle)."""


def oddeven_merge_sort(n):
    pairs = []
    p = 1
    while p < n:
        k = p
        while k >= 1:
            for j in range(k % p, n - k, 2 * k):
                for i in range(min(k, n - j - k)):
                    if (i + j) // (2 * p) == (i + j + k) // (2 * p):
                        pairs.append((i + j, i + j + k))
            k //= 2
        p *= 2
    return pairs


if __name__ == "__main__":
    pairs = oddeven_merge_sort(16)
    for m in range(1 << 16):
        v = [(m >> i) & 1 for i in range(16)]
        for a, b in pairs:
            if v[a] > v[b]:
                v[a], v[b] = v[b], v[a]
        assert all(v[i] <= v[i + 1] for i in range(15)), m
    print(len(pairs), "compare-exchanges; sorts every 0-1 input")
    print(", ".join("{%d,%d}" % p for p in pairs))
