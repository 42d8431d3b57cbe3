"""Host-step variants on the bare stepper (wall clock per synchronous step):
   A. zbot_step_host, pipelined DMA ranges;  B. zbot_step with pinned-host pointers (pure zero-copy);
   C. zero-copy actions in, device outputs + one packed D2H."""
import ctypes as C, sys, os, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zbot_lab_b200.stepper import NativeStepper, _stream
from zbot_lab_b200 import native

def wall(fn, reps=100):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6

for n in [int(x) for x in sys.argv[1:]] or [4096, 16384, 65536]:
    st = NativeStepper(n, "cuda:0")
    st.reset_idx(None)
    h_act = torch.randn(n, 6).pin_memory()
    h_rows = torch.empty(n, 25).pin_memory()
    h_packed = torch.empty(n * 98, dtype=torch.uint8).pin_memory()
    print(n, f"A zbot_step_host (packed rows, zero-copy): {wall(lambda: st.step_host(h_act, h_rows)):7.1f} us", flush=True)
    base = h_packed.data_ptr()
    vp = C.c_void_p
    def zc():
        slot, prev = st._next_slot()
        native.check(st.lib.zbot_step(st._h, vp(h_act.data_ptr()), vp(base), vp(base + n * 92), vp(base + n * 96),
                                      vp(base + n * 97), slot, prev, _stream(st.device)))
        torch.cuda.current_stream().synchronize()
    print(n, f"B zero-copy in+out:          {wall(zc):7.1f} us", flush=True)
    d_packed = torch.empty(n * 98, dtype=torch.uint8, device="cuda:0")
    db = d_packed.data_ptr()
    def zc_in():
        slot, prev = st._next_slot()
        native.check(st.lib.zbot_step(st._h, vp(h_act.data_ptr()), vp(db), vp(db + n * 92), vp(db + n * 96),
                                      vp(db + n * 97), slot, prev, _stream(st.device)))
        h_packed.copy_(d_packed, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    print(n, f"C zero-copy in, DMA out:     {wall(zc_in):7.1f} us", flush=True)
    d_act = torch.empty(n, 6, device="cuda:0")
    def zc_out():
        d_act.copy_(h_act, non_blocking=True)
        slot, prev = st._next_slot()
        native.check(st.lib.zbot_step(st._h, vp(d_act.data_ptr()), vp(base), vp(base + n * 92), vp(base + n * 96),
                                      vp(base + n * 97), slot, prev, _stream(st.device)))
        torch.cuda.current_stream().synchronize()
    print(n, f"D DMA in, zero-copy out:     {wall(zc_out):7.1f} us", flush=True)
    st.close()
