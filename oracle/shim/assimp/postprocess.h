// Stand-in for <assimp/postprocess.h> (no post-processing flags are used: ReadFile(path, 0)).
#pragma once
