// hostcopy_check.cpp -- CPU test of the pageable path's byte mover (csrc/raisr_hostcopy.cpp): random sizes, alignments and row
// strides against memcpy(), guard bytes around every destination row.  Built and run by tests/test_abi.py.
#include "raisr_hostcopy.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

int main()
{
    std::srand(7);
    std::printf("mode %d\n", raisr::host_copy_mode());
    for (int t = 0; t < 400; ++t) {
        const size_t row = (t % 5 == 0) ? (size_t)(std::rand() % 200) : (size_t)(std::rand() % 70000);
        const int rows = 1 + std::rand() % 5;
        const bool packed = t % 3 == 0;
        const size_t sstep = packed ? row : row + (size_t)(std::rand() % 130), dstep = packed ? row : row + (size_t)(std::rand() % 130);
        const size_t so = (size_t)(std::rand() % 64), dofs = 64 + (size_t)(std::rand() % 64);
        std::vector<unsigned char> src(so + sstep * rows + 64), dst(dofs + dstep * rows + 128, 0xAB), want;
        for (auto &b : src) b = (unsigned char)std::rand();
        want = dst;
        for (int y = 0; y < rows; ++y) std::memcpy(want.data() + dofs + y * dstep, src.data() + so + y * sstep, row);
        raisr::host_copy_rows(dst.data() + dofs, dstep, src.data() + so, sstep, row, rows);
        if (dst != want) { std::printf("mismatch: row %zu rows %d sstep %zu dstep %zu\n", row, rows, sstep, dstep); return 1; }
    }
    std::printf("ok\n");
    return 0;
}
