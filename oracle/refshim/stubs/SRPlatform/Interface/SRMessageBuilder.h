#pragma once
#include "../SRPlatform/Interface/SRException.h"
