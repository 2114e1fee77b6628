#include "../ffstub.h"
