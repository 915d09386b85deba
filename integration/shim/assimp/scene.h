// empty on purpose: types live in assimp/Importer.hpp (host oracle shim)
#pragma once
