// Stand-in for <assimp/Importer.hpp>: Assimp::Importer::ReadFile over the in-repo .assbin reader.
#pragma once
#include "scene.h"

namespace Assimp {
class Importer {
    aiScene* m_scene = nullptr;
    static aiNode* convertNode(const slr::assbin::Node& n) {
        aiNode* o = new aiNode();
        o->mName.s = n.name;
        std::memcpy(&o->mTransformation, n.transform, sizeof(float) * 16);
        o->meshStorage.assign(n.meshes.begin(), n.meshes.end());
        o->mNumMeshes = (unsigned)o->meshStorage.size(); o->mMeshes = o->meshStorage.data();
        for (const slr::assbin::Node& c : n.children) o->childStorage.push_back(convertNode(c));
        o->mNumChildren = (unsigned)o->childStorage.size(); o->mChildren = o->childStorage.data();
        return o;
    }
    static void fill(std::vector<aiVector3D>& dst, const std::vector<float>& src) {
        dst.resize(src.size() / 3);
        for (size_t i = 0; i < dst.size(); ++i) dst[i] = aiVector3D(src[3 * i], src[3 * i + 1], src[3 * i + 2]);
    }
public:
    ~Importer() { delete m_scene; }
    const aiScene* ReadFile(const std::string& path, unsigned int) {
        slr::assbin::Scene src;
        std::string err;
        if (!slr::assbin::load(path, &src, &err)) { fprintf(stderr, "%s\n", err.c_str()); return nullptr; }
        delete m_scene;
        m_scene = new aiScene();
        for (const slr::assbin::Mesh& m : src.meshes) {
            aiMesh* o = new aiMesh();
            std::memset(o->mTextureCoords, 0, sizeof(o->mTextureCoords));
            std::memset(o->mNumUVComponents, 0, sizeof(o->mNumUVComponents));
            o->mPrimitiveTypes = m.primitiveTypes; o->mNumVertices = m.numVertices(); o->mMaterialIndex = m.materialIndex;
            fill(o->storage[0], m.positions); fill(o->storage[1], m.normals); fill(o->storage[2], m.tangents);
            fill(o->storage[3], m.bitangents); fill(o->storage[4], m.texCoords);
            o->mVertices = o->storage[0].data();
            o->mNormals = o->storage[1].empty() ? nullptr : o->storage[1].data();
            o->mTangents = o->storage[2].empty() ? nullptr : o->storage[2].data();
            o->mBitangents = o->storage[3].empty() ? nullptr : o->storage[3].data();
            o->mTextureCoords[0] = o->storage[4].empty() ? nullptr : o->storage[4].data();
            o->mNumUVComponents[0] = o->storage[4].empty() ? 0 : m.numUVComponents;
            o->indexStorage = m.indices;
            o->faces.resize(m.indices.size() / 3);
            for (size_t f = 0; f < o->faces.size(); ++f) { o->faces[f].mNumIndices = 3; o->faces[f].mIndices = &o->indexStorage[3 * f]; }
            o->mNumFaces = (unsigned)o->faces.size(); o->mFaces = o->faces.data();
            o->mName.s = m.name;
            m_scene->meshStorage.push_back(o);
        }
        for (const slr::assbin::Material& m : src.materials) { aiMaterial* o = new aiMaterial(); o->data = m; m_scene->materialStorage.push_back(o); }
        m_scene->mNumMeshes = (unsigned)m_scene->meshStorage.size(); m_scene->mMeshes = m_scene->meshStorage.data();
        m_scene->mNumMaterials = (unsigned)m_scene->materialStorage.size(); m_scene->mMaterials = m_scene->materialStorage.data();
        m_scene->mRootNode = convertNode(src.root);
        return m_scene;
    }
};
}
