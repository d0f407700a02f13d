#pragma once
#include "../simt_emu.h"
