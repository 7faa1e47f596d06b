#include "qtshim.h"
