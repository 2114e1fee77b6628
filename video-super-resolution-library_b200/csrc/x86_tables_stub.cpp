#include "x86_tables.h"
namespace raisr {
__attribute__((weak)) bool x86_tables(X86Tables *t)
{
    t->rsqrt14 = t->rcp14 = nullptr;
    t->rsqrtps = t->rcpps = nullptr;
    return false;
}
}  // namespace raisr
