// RtModelAssimp.cpp — the reference's mesh loader (libs/DXRFramework/RtModel.cpp:26-58) for hosts that have Assimp.
//
// Compiled only with RT_HAVE_ASSIMP (make ASSIMP_INCLUDE=<dir holding assimp/> [ASSIMP_LIB="-L... -lassimp"]); the
// reference vendors Assimp's headers and Windows binaries only, so the default build leaves this file out and RtModel
// falls back to its OBJ reader.  Behaviour kept from the reference: the post-processing flags, every mesh of the scene
// merged into ONE interleaved {position, normal} buffer with a per-mesh vertex offset added to the indices, zero
// normals when a mesh has none, and three indices per face (aiProcess_Triangulate guarantees triangles; other faces
// are skipped instead of asserting).
#if defined(RT_HAVE_ASSIMP)
#include <assimp/cimport.h>
#include <assimp/postprocess.h>
#include <assimp/scene.h>

#include "RtModel.h"

namespace DXRFramework {

bool RtModel::loadWithAssimp(const std::string &path, std::vector<Vertex> &vertices, std::vector<uint32_t> &indices) {
    const unsigned flags = aiProcess_Triangulate | aiProcess_GenSmoothNormals | aiProcess_FlipUVs | aiProcess_JoinIdenticalVertices |
                           aiProcess_PreTransformVertices;
    const aiScene *scene = aiImportFile(path.c_str(), flags);
    if (!scene) return false;
    vertices.clear();
    indices.clear();
    uint32_t base = 0;
    for (unsigned meshId = 0; meshId < scene->mNumMeshes; ++meshId) {
        const aiMesh *mesh = scene->mMeshes[meshId];
        const bool hasNormals = mesh->HasNormals();
        for (unsigned i = 0; i < mesh->mNumVertices; ++i) {
            const aiVector3D &p = mesh->mVertices[i];
            Vertex v{};
            v.position = DirectX::XMFLOAT3{p.x, p.y, p.z};
            v.normal = hasNormals ? DirectX::XMFLOAT3{mesh->mNormals[i].x, mesh->mNormals[i].y, mesh->mNormals[i].z} : DirectX::XMFLOAT3{0.0f, 0.0f, 0.0f};
            vertices.push_back(v);
        }
        for (unsigned i = 0; i < mesh->mNumFaces; ++i) {
            const aiFace &face = mesh->mFaces[i];
            if (face.mNumIndices != 3) continue;
            indices.push_back(base + face.mIndices[0]);
            indices.push_back(base + face.mIndices[1]);
            indices.push_back(base + face.mIndices[2]);
        }
        base += mesh->mNumVertices;
    }
    aiReleaseImport(scene);
    return !vertices.empty() && !indices.empty();
}

}  // namespace DXRFramework
#endif
