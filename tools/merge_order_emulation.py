"""Host emulation of the slot walk of rs_merge_tiles_pipelined_kernel<.., REGIONS = true> (seek / advance / locate as in csrc/rs_merge.cuh,
order as in csrc/rs_merge_order.h): every tile of every job is visited exactly once for any job sizes and grid, and
the work per CTA is balanced even when all tiles sit in one job.  Run by tests/test_capi_load.py (no GPU needed)."""
import random

R = 64


def emulate(job_tiles, resident):
    starts = [0]
    for t in job_tiles:
        starts.append(starts[-1] + t)
    T, J = starts[-1], len(job_tiles)
    G = min(resident, T)                      # the host launches min(tiles, resident CTAs); the stride is gridDim.x
    first = [starts[j] if j < J else T for j in range(8)]
    C = (T + R - 1) // R
    S = C * R

    def tile(p):
        return (p % R) * C + p // R

    def seek(s):
        while s < S:
            if tile(s) < T:
                break
            if S - s <= G:
                return S
            s += G
        return s

    def advance(s):
        return S if S - s <= G else seek(s + G)

    def locate(s):
        t, job = tile(s), 0
        for j in range(1, 8):
            if t >= first[j]:
                job = j
        return job, t - first[job]

    seen, per_cta = set(), []
    for c in range(G):
        s, n = seek(c), 0
        while s < S:
            job, lt = locate(s)
            assert job < J and 0 <= lt < job_tiles[job], (job_tiles, resident, c, s, job, lt)
            assert (job, lt) not in seen, (job_tiles, resident, job, lt)
            seen.add((job, lt))
            n += 1
            s = advance(s)
        per_cta.append(n)
    assert len(seen) == T, (job_tiles, resident, len(seen), T)
    return per_cta


def main():
    random.seed(1)
    cases = [[5], [0, 5], [5, 0], [0, 0, 3, 0], [1] * 8, [1000, 0, 0, 0], [0, 0, 0, 1000], [300, 0, 2],
             [733, 1, 0, 555, 64, 63, 65, 2], [26214, 26214, 26217, 26210]]
    for _ in range(40):
        cases.append([random.choice([0, 1, 2, 17, 300, 4096]) for _ in range(random.randint(1, 8))])
    n = 0
    for jt in cases:
        if sum(jt) == 0:
            continue
        for resident in (1, 7, 64, 148, 592, 888):
            emulate(jt, resident)
            n += 1
    skew = emulate([104858, 0, 0, 0], 592)
    even = emulate([26214, 26214, 26217, 26210], 592)
    assert max(skew) - min(skew) <= 4 and max(even) - min(even) <= 4, (min(skew), max(skew), min(even), max(even))
    print(f"{n} configurations: every tile exactly once; tiles per CTA skewed batch {min(skew)}..{max(skew)}, "
          f"balanced batch {min(even)}..{max(even)}")


if __name__ == "__main__":
    main()
