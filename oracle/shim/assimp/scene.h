// Stand-in for <assimp/scene.h>: just the types and accessors libSLRSceneGraph touches
// (node_constructor.cpp:35-178, API.cpp:800-925), backed by the in-repo .assbin reader.
// Test infrastructure: lets the reference's own scene loader run in an image without assimp.
#pragma once
#include <cstring>
#include <string>
#include <vector>
#include "../../../slr_b200/host/assets/assbin.h"

enum aiReturn { aiReturn_SUCCESS = 0, aiReturn_FAILURE = -1 };
enum aiPrimitiveType { aiPrimitiveType_POINT = 1, aiPrimitiveType_LINE = 2, aiPrimitiveType_TRIANGLE = 4, aiPrimitiveType_POLYGON = 8 };
enum aiTextureType {
    aiTextureType_NONE = 0, aiTextureType_DIFFUSE = 1, aiTextureType_SPECULAR = 2, aiTextureType_AMBIENT = 3, aiTextureType_EMISSIVE = 4,
    aiTextureType_HEIGHT = 5, aiTextureType_NORMALS = 6, aiTextureType_SHININESS = 7, aiTextureType_OPACITY = 8, aiTextureType_DISPLACEMENT = 9
};

struct aiString {
    std::string s;
    const char* C_Str() const { return s.c_str(); }
};
struct aiVector3D {
    float x, y, z;
    aiVector3D() : x(0), y(0), z(0) {}
    aiVector3D(float a, float b, float c) : x(a), y(b), z(c) {}
};
struct aiMatrix4x4 { float a1, a2, a3, a4, b1, b2, b3, b4, c1, c2, c3, c4, d1, d2, d3, d4; };
struct aiFace { unsigned int mNumIndices; unsigned int* mIndices; };

struct aiMesh {
    unsigned int mPrimitiveTypes, mNumVertices, mNumFaces, mMaterialIndex;
    aiVector3D* mVertices; aiVector3D* mNormals; aiVector3D* mTangents; aiVector3D* mBitangents;
    aiVector3D* mTextureCoords[8];
    unsigned int mNumUVComponents[8];
    aiFace* mFaces;
    aiString mName;
    std::vector<aiVector3D> storage[5];
    std::vector<aiFace> faces;
    std::vector<unsigned int> indexStorage;
};

struct aiNode {
    aiString mName;
    aiMatrix4x4 mTransformation;
    unsigned int mNumMeshes; unsigned int* mMeshes;
    unsigned int mNumChildren; aiNode** mChildren;
    std::vector<unsigned int> meshStorage;
    std::vector<aiNode*> childStorage;
    ~aiNode() { for (aiNode* c : childStorage) delete c; }
};

// material keys: (key string, semantic, index)
#define AI_MATKEY_NAME "?mat.name", 0, 0
#define AI_MATKEY_COLOR_DIFFUSE "$clr.diffuse", 0, 0
#define AI_MATKEY_COLOR_SPECULAR "$clr.specular", 0, 0
#define AI_MATKEY_COLOR_EMISSIVE "$clr.emissive", 0, 0
#define AI_MATKEY_TEXTURE(type, N) "$tex.file", type, N
#define AI_MATKEY_TEXTURE_DIFFUSE(N) AI_MATKEY_TEXTURE(aiTextureType_DIFFUSE, N)
#define AI_MATKEY_TEXTURE_SPECULAR(N) AI_MATKEY_TEXTURE(aiTextureType_SPECULAR, N)
#define AI_MATKEY_TEXTURE_EMISSIVE(N) AI_MATKEY_TEXTURE(aiTextureType_EMISSIVE, N)
#define AI_MATKEY_TEXTURE_HEIGHT(N) AI_MATKEY_TEXTURE(aiTextureType_HEIGHT, N)
#define AI_MATKEY_TEXTURE_NORMALS(N) AI_MATKEY_TEXTURE(aiTextureType_NORMALS, N)
#define AI_MATKEY_TEXTURE_OPACITY(N) AI_MATKEY_TEXTURE(aiTextureType_OPACITY, N)
#define AI_MATKEY_TEXTURE_DISPLACEMENT(N) AI_MATKEY_TEXTURE(aiTextureType_DISPLACEMENT, N)

struct aiMaterial {
    slr::assbin::Material data;
    aiReturn Get(const char* key, unsigned int type, unsigned int idx, aiString& out) const {
        return data.getString(key, type, idx, &out.s) ? aiReturn_SUCCESS : aiReturn_FAILURE;
    }
    aiReturn Get(const char* key, unsigned int, unsigned int, float* out, unsigned int*) const {
        return data.getColor(key, out) ? aiReturn_SUCCESS : aiReturn_FAILURE;
    }
    unsigned int GetTextureCount(aiTextureType type) const { return data.textureCount((uint32_t)type); }
};

struct aiScene {
    unsigned int mNumMeshes; aiMesh** mMeshes;
    unsigned int mNumMaterials; aiMaterial** mMaterials;
    aiNode* mRootNode;
    std::vector<aiMesh*> meshStorage;
    std::vector<aiMaterial*> materialStorage;
    ~aiScene() { for (aiMesh* m : meshStorage) delete m; for (aiMaterial* m : materialStorage) delete m; delete mRootNode; }
};
