"""The device grid builder sorts every cell's list in one thread (rg_grid.cu: gb_sort_u32, insertion
sort up to 32 items, heap sort beyond).  The routine is __host__ __device__: compile its text with
g++ and check it against std::sort, under ASan/UBSan.  CPU-only."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r'''
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>
%s
int main() {
    std::mt19937 rng(7);
    long checked = 0;
    for (uint32_t n : {0u, 1u, 2u, 3u, 5u, 31u, 32u, 33u, 34u, 63u, 64u, 100u, 257u, 1000u, 4097u, 50000u})
        for (int rep = 0; rep < (n < 2000 ? 200 : 3); ++rep) {
            std::vector<uint32_t> a(n);
            const int mode = rep %% 5;   // random, ascending, descending, few distinct values, many ties
            for (uint32_t i = 0; i < n; ++i)
                a[i] = mode == 0 ? (uint32_t)rng() : mode == 1 ? i : mode == 2 ? n - i : mode == 3 ? (uint32_t)(rng() %% 4) : (uint32_t)(rng() %% (n + 1));
            std::vector<uint32_t> b = a;
            std::sort(b.begin(), b.end());
            gb_sort_u32(a.data(), n);
            if (a != b) { std::printf("MISMATCH n=%%u mode=%%d\n", n, mode); return 1; }
            ++checked;
        }
    std::printf("ok %%ld\n", checked);
    return 0;
}
'''


def test_cell_list_sort_matches_std_sort(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    src = open(os.path.join(ROOT, "raingun_b200", "csrc", "rg_grid.cu")).read()
    m = re.search(r"__host__ __device__ inline void gb_sort_u32\(uint32_t \*v, uint32_t n\) \{.*?\n\}\n", src, re.S)
    assert m, "gb_sort_u32 not found in rg_grid.cu"
    cpp = tmp_path / "sort_test.cpp"
    cpp.write_text(HARNESS % m.group(0).replace("__host__ __device__ ", ""))
    exe = tmp_path / "sort_test"
    r = subprocess.run([gxx, "-std=c++17", "-O2", "-fsanitize=address,undefined", str(cpp), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
