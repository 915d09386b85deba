// empty on purpose: see cuda_runtime_api.h in this directory (host oracle shim)
#pragma once
