import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import subprocess
# the probe is a tool, not part of the product library: built on first use into tools/_build/ (nvcc cross-compiles sm_100a)
SRC, OUT = os.path.join(ROOT, "tools", "tma_probe.cu"), os.path.join(ROOT, "tools", "_build", "libtma_probe.so")
CSRC = os.path.join(ROOT, "seamless_communication_b200", "csrc")
if not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-shared",
                           "-Xcompiler", "-fPIC", "-I", CSRC, "-I", os.path.join(ROOT, "include"), SRC,
                           os.path.join(CSRC, "library.cu"), "-o", OUT, "-lcuda"])
lib = C.CDLL(OUT)
lib.sb_tma_probe.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
rows, ld = 1 << 18, 1024  # 512 MB fp16
buf = torch.randn(rows, ld, device="cuda").half()
iters = 512
print("ctas stages box_rows issuers -> B/clk per CTA (avg), total GB/s")
for mode in (0,):
  print('wait mode', mode, '(0 = try_wait, 1 = test_wait spin)')
  for ctas in (1, 148):
      for stages, box_rows, issuers in [(2, 128, 1), (4, 128, 1), (8, 128, 1), (4, 256, 1), (4, 64, 1), (4, 128, 4)]:
          cyc = torch.zeros(ctas, dtype=torch.int64, device="cuda")
          e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          rc = lib.sb_tma_probe(buf.data_ptr(), rows, ld, ctas, stages, box_rows, issuers, iters, cyc.data_ptr(), None, mode)
          torch.cuda.synchronize()
          e0.record()
          rc = lib.sb_tma_probe(buf.data_ptr(), rows, ld, ctas, stages, box_rows, issuers, iters, cyc.data_ptr(), None, mode)
          e1.record(); torch.cuda.synchronize()
          assert rc == 0
          ms = e0.elapsed_time(e1)
          bytes_cta = iters * box_rows * 128
          print(f"{ctas:4d} {stages:2d} {box_rows:4d} {issuers:2d} -> {bytes_cta / cyc.float().mean().item():6.1f} B/clk/CTA   {ctas * bytes_cta / ms / 1e6:8.0f} GB/s  ({ms*1e3:.0f} us)")
  