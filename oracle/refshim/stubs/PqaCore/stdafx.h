// TEST INFRASTRUCTURE stub of PqaCore/stdafx.h: same role (pull in SRPlatform), minus <windows.h>/<cuda_runtime.h>.
#pragma once
#include "../SRPlatform/Interface/ISRLogger.h"
#include "../SRPlatform/Interface/SRAccumulator.h"
#include "../SRPlatform/Interface/SRAccumVectDbl256.h"
#include "../SRPlatform/Interface/SRBaseSubtask.h"
#include "../SRPlatform/Interface/SRBaseTask.h"
#include "../SRPlatform/Interface/SRBasicTypes.h"
#include "../SRPlatform/Interface/SRBitArray.h"
#include "../SRPlatform/Interface/SRCast.h"
#include "../SRPlatform/Interface/SRCpuInfo.h"
#include "../SRPlatform/Interface/SRDoubleNumber.h"
#include "../SRPlatform/Interface/SRException.h"
#include "../SRPlatform/Interface/SRHeap.h"
#include "../SRPlatform/Interface/SRLogStream.h"
#include "../SRPlatform/Interface/SRMath.h"
#include "../SRPlatform/Interface/SRMaxSizeof.h"
#include "../SRPlatform/Interface/SRMemPool.h"
#include "../SRPlatform/Interface/SRStandardSubtask.h"
#include "../SRPlatform/Interface/SRThreadPool.h"
#include "../SRPlatform/Interface/SRUtils.h"
#include "../SRPlatform/Interface/SRPoolRunner.h"
#include "../SRPlatform/Interface/SRSimd.h"
#include "../SRPlatform/Interface/SRVectMath.h"
// MSVC's lax two-phase lookup lets ProbQA headers name SRPlat members unqualified (e.g. SRMath in
// CEEvalQsSubtaskConsider.h:17); make the same names visible for g++.
namespace ProbQA { using namespace SRPlat; }
// Some reference headers rely on the precompiled header's include order (e.g. CEHeapifyPriorsTask.h uses
// CEBaseTask without including it).
#include "../PqaCore/CEBaseTask.h"
