"""Fluid model of the trixib200_rhs_host pipeline (upload -> compute -> download over chunks of the Morton order) at
level 7: PCIe rates from profiles/r1_pcie_probe.json (55.5 / 57.2 GB/s one direction alone, 50 GB/s each when both run),
compute ignored. A chunk can be downloaded once the chunks of all its face neighbours (periodic) have been uploaded.

    python tools/host_pipe_model.py

Reproduces the measured 139 ms per rhs! for the default 16-layer chunks (model 138.6 ms for 8 z-slabs) and shows what the
upload order and the slab thickness are worth (DESIGN.md sections 5 and 8)."""
N = 8                                   # chunks per dimension at 4096 elements (16^3) per chunk, level 7
GB = 5.36870912                         # one state vector


def deinterleave(i):
    x = y = z = 0
    for b in range(3):
        x |= ((i >> (3 * b)) & 1) << b
        y |= ((i >> (3 * b + 1)) & 1) << b
        z |= ((i >> (3 * b + 2)) & 1) << b
    return x, y, z


def interleave(x, y, z):
    i = 0
    for b in range(3):
        i |= ((x >> b) & 1) << (3 * b) | ((y >> b) & 1) << (3 * b + 1) | ((z >> b) & 1) << (3 * b + 2)
    return i


def run(n_units, unit_gb, ready_after):
    """ready_after[u] = index (in upload order) of the upload after which unit u may be downloaded."""
    by_upload = [[] for _ in range(n_units)]
    for u, r in enumerate(ready_after):
        by_upload[r].append(u)
    t, up_i, up_left, queue, dl_left, done, first = 0.0, 0, unit_gb, [], 0.0, 0, None
    while done < n_units:
        if dl_left <= 0 and queue:
            queue.pop(0)
            dl_left = unit_gb
            first = t if first is None else first
        ua, da = up_i < n_units, dl_left > 0
        ru = (50.0 if da else 55.5) if ua else 0.0
        rd = (50.0 if ua else 57.2) if da else 0.0
        dt = min(([up_left / ru] if ua else []) + ([dl_left / rd] if da else []))
        t += dt
        if ua:
            up_left -= ru * dt
            if up_left <= 1e-15:
                queue.extend(by_upload[up_i]); up_i += 1; up_left = unit_gb
        if da:
            dl_left -= rd * dt
            if dl_left <= 1e-15:
                dl_left = 0.0; done += 1
    return t * 1e3, first * 1e3


def chunk_orders():
    nch = N ** 3
    nbrs = []
    for c in range(nch):
        x, y, z = deinterleave(c)
        nbrs.append({interleave((x + a) % N, (y + b) % N, (z + d) % N)
                     for a, b, d in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1))} | {c})
    orders = {"z-layer order (the library's)": sorted(range(nch), key=lambda c: (deinterleave(c)[2], c)),
              "Morton order": list(range(nch)),
              "diagonal wavefront": sorted(range(nch), key=lambda c: (sum(deinterleave(c)), c))}
    for name, order in orders.items():
        upos = {c: i for i, c in enumerate(order)}
        total, first = run(nch, GB / nch, [max(upos[n] for n in nbrs[c]) for c in range(nch)])
        print(f"{name:32s} {total:6.1f} ms per rhs!, first download after {first:5.1f} ms")


def slabs():
    for nz in (8, 16, 32, 64):
        total, _ = run(nz, GB / nz, [max((s - 1) % nz, s, (s + 1) % nz) for s in range(nz)])
        print(f"{nz:3d} z-slabs, slab granularity: {total:6.1f} ms per rhs!")


if __name__ == "__main__":
    chunk_orders()
    slabs()
    print("lower bound (both directions streaming all the time): %.1f ms" % (GB / 50.0 * 1e3))
