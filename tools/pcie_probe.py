"""PCIe copy bandwidth of the box: pinned host <-> device, one direction and both at once (what bounds e2e)."""
import json, torch
n = 1 << 30  # 1 GiB
h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, reps=4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if down:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
run(True, True, 1)
print(json.dumps({"h2d_gbs": run(True, False), "d2h_gbs": run(False, True), "duplex_each_gbs": run(True, True)}))
