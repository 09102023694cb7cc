#include "petscdm.h"
