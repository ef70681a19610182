#pragma once
// compat.h is force-included; nothing else needed.
