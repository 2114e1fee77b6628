#include "x86_tables.h"
namespace raisr {
__attribute__((weak)) void x86_tables(const uint16_t *src[4], size_t n[4])
{
    for (int i = 0; i < 4; ++i) { src[i] = nullptr; n[i] = 0; }
}
}
