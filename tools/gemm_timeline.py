import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seamless_communication_b200 import _lib, ops
from seamless_communication_b200.ops import Seq
lib = _lib.load()
raw = C.CDLL(_lib.LIB_PATH)
raw.sb_gemm_debug_timeline.argtypes = [C.c_void_p]
tl = torch.zeros(48, dtype=torch.int64, device="cuda")
raw.sb_gemm_debug_timeline(tl.data_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(name, fn):
    for _ in range(3): fn()
    res = []
    for _ in range(5):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t = tl.cpu().tolist()
        res.append(([t[i] - t[0] for i in range(1, 7)], e0.elapsed_time(e1) * 1e3))
    print(name)
    for r, ev in res[1:]:
        if r is res[-1][0]:
            t = tl.cpu().tolist()
            print("   per k-block (ns from entry) [before wait, after wait, mmas issued, after commit]:")
            print("   " + " | ".join("%d:%d,%d,%d,%d" % (k, *[t[8 + 4 * k + j] - t[0] for j in range(4)]) for k in range(8)))
        print("   ns since kernel entry: setup %5d | first full %5d | mma issued %5d | tmem_full seen %5d | epilogue done %5d | after sync %5d   (event %.1f us)" % (*r, ev))
x = Seq(1, 160, 1024); x.buf.normal_()
w = (torch.randn(1024, 1024, device="cuda") * 0.02).half(); b = torch.randn(1024, device="cuda")
out = Seq(1, 160, 1024)
run("dec out-proj 160x1024x1024 (BN=64, 16 k-blocks)", lambda: ops.gemm(x, w, 1024, b, out=out))
part = torch.empty(4 * 256, 1024, dtype=torch.float32, device="cuda")
run("split-K 4 (BN=128, 4 k-blocks)", lambda: ops.gemm_splitk(x, w, 1024, 4, part))
x8 = Seq(1, 160, 8192); x8.buf.normal_(); w8 = (torch.randn(1024, 8192, device="cuda") * 0.02).half()
part8 = torch.empty(8 * 256, 1024, dtype=torch.float32, device="cuda")
run("split-K 8 ffn2 (16 k-blocks)", lambda: ops.gemm_splitk(x8, w8, 1024, 8, part8))
xe = Seq(1, 32 * 499, 1024); xe.buf.normal_(); we = (torch.randn(4096, 1024, device="cuda") * 0.02).half(); be = torch.randn(4096, device="cuda")
oe = Seq(1, 32 * 499, 4096)
run("enc ffn1 tile (0,0)", lambda: ops.gemm(xe, we, 4096, be, act=ops.ACT_SILU, out=oe))
