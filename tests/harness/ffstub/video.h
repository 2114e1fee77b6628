#include "ffstub.h"
