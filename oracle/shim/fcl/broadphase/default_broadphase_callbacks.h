#pragma once
#include "../fcl_stub.h"
