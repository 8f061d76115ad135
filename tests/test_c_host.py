"""The drop-in boundary is a plain C ABI: include/b2z.h (+ protocol_params.h) must compile as C99 with no C++ or torch types,
a C translation unit must link against libb2z.so, run without a GPU and FAIL LOUDLY there (no CPU fallback), and the C++
example host (tools/microbench/session_lat.cpp) must build against the same header."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "zeth_b200")

C_HOST = r"""
#include <stdio.h>
#include <string.h>
#include "b2z.h"
#include "protocol_params.h"

int main(void) {
  b2z_ctx* ctx = NULL;
  b2z_err e;
  printf("version=%s\n", b2z_version());
  printf("queries=%d inv_rate=%d top32=%d\n", (int)PP_QUERIES, (int)PP_INV_RATE, (int)((size_t)1 << pp_merkle_top_layer(22, PP_QUERIES)));
  e = b2z_init(0, &ctx);
  if (e) { printf("init_error=%s\n", e); return ctx == NULL ? 0 : 3; }
  b2z_destroy(ctx);
  printf("init_ok\n");
  return 0;
}
"""


def _have_lib():
    from zeth_b200 import build as b
    b.build()
    return os.path.exists(os.path.join(LIBDIR, "libb2z.so"))


def test_header_is_plain_c99_and_a_c_host_links_and_fails_loudly_without_gpu():
    assert _have_lib()
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "host.c"), os.path.join(d, "host")
        open(src, "w").write(C_HOST)
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-o", exe, src,
                               "-L", LIBDIR, "-lb2z", f"-Wl,-rpath,{LIBDIR}"])
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "version=b2z" in out.stdout
        import torch
        if not torch.cuda.is_available():
            # no device: the library says so instead of computing on the CPU
            assert "init_error=" in out.stdout and "no CPU fallback" in out.stdout, out.stdout


def test_cpp_example_host_builds_against_the_header():
    assert _have_lib()
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "session_lat")
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tools", "microbench", "session_lat.cpp"),
                               "-L", LIBDIR, "-lb2z", f"-Wl,-rpath,{LIBDIR}"])
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        assert out.returncode == 2 and "usage" in out.stderr
