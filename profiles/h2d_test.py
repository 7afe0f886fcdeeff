import torch, time
n = 384**3
a = torch.empty(n, dtype=torch.float32).pin_memory(); b = torch.empty(n, dtype=torch.float32).pin_memory()
da = torch.empty(n, dtype=torch.float32, device='cuda'); db = torch.empty_like(da)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def seq():
    da.copy_(a, non_blocking=True); db.copy_(b, non_blocking=True)
def par():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): da.copy_(a, non_blocking=True)
    with torch.cuda.stream(s2): db.copy_(b, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
def chunks(k=8):
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    step = n // k
    for i in range(k):
        with torch.cuda.stream(s1 if i % 2 == 0 else s2):
            da[i*step:(i+1)*step].copy_(a[i*step:(i+1)*step], non_blocking=True)
            db[i*step:(i+1)*step].copy_(b[i*step:(i+1)*step], non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
for name, fn in (('sequential', seq), ('two streams', par), ('chunked 2 streams', chunks)):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print('%-20s %.2f ms  %.1f GB/s' % (name, ms, 2 * n * 4 / ms / 1e6))
