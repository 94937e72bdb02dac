"""Stub (test infrastructure): python-Levenshtein is absent from this image; the reference uses Levenshtein.distance for the dev-set edit
distance (src/train.py:405-419, src/utils.py).  Unit-cost edit distance, same definition."""


def distance(a, b):
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]
