// Minimal stand-in for the part of Assimp the reference calls (Scene.cpp:226-291), used wherever the reference's own Scene.cpp is
// compiled in this image: integration/build_in_reference_tree.sh (the drop-in proof) and the oracle builds (oracle/build_ref*.sh include
// it from here; nothing under integration/ depends on oracle/).  Assimp is not vendored by the reference and is absent here, so
// its behaviour for the bundled Wavefront files is restated: ReadFile() without
// aiProcess_JoinIdenticalVertices emits ONE vertex per face corner, in file order, as a
// single mesh hanging off the root node.  Parity unpinned for Assimp
// itself (no Assimp build exists to compare with); pinned only by the committed Render.bmp.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct aiVector3D { float x, y, z; };
struct aiFace { unsigned int mNumIndices; unsigned int* mIndices; };
struct aiMesh {
    unsigned int mNumVertices; aiVector3D* mVertices; aiVector3D* mNormals;
    unsigned int mNumFaces; aiFace* mFaces;
};
struct aiNode { unsigned int mNumMeshes; unsigned int* mMeshes; unsigned int mNumChildren; aiNode** mChildren; };
struct aiScene { unsigned int mFlags; aiNode* mRootNode; aiMesh** mMeshes; };
enum { AI_SCENE_FLAGS_INCOMPLETE = 1 };
enum { aiProcess_FlipUVs = 0x800000 };

namespace Assimp {
class Importer {
    std::vector<aiVector3D> pos_, nrm_;
    std::vector<unsigned int> idx_;
    std::vector<aiFace> faces_;
    aiMesh mesh_{}; aiMesh* meshes_[1]; aiNode root_{}; unsigned int root_mesh_[1]; aiScene scene_{};
    std::string err_;
public:
    const char* GetErrorString() const { return err_.c_str(); }
    const aiScene* ReadFile(const std::string& path_in, unsigned int)
    {
        std::string path = path_in;
        for (char& c : path) if (c == '\\') c = '/';
        FILE* f = fopen(path.c_str(), "r");
        if (!f) { err_ = "Unable to open file \"" + path_in + "\"."; return nullptr; }
        std::vector<aiVector3D> v, vn;
        char line[1024];
        while (fgets(line, sizeof line, f)) {
            if (line[0] == 'v' && line[1] == ' ') {
                aiVector3D p; sscanf(line + 2, "%f %f %f", &p.x, &p.y, &p.z); v.push_back(p);
            } else if (line[0] == 'v' && line[1] == 'n') {
                aiVector3D p; sscanf(line + 3, "%f %f %f", &p.x, &p.y, &p.z); vn.push_back(p);
            } else if (line[0] == 'f' && line[1] == ' ') {
                int n = 0; char* s = line + 2;
                unsigned int first = (unsigned)pos_.size();
                while (*s) {
                    while (*s == ' ' || *s == '\t') ++s;
                    if (*s == '\0' || *s == '\n' || *s == '\r') break;
                    long iv = strtol(s, &s, 10), it = 0, in = 0;
                    if (*s == '/') { ++s; if (*s != '/') it = strtol(s, &s, 10); if (*s == '/') { ++s; in = strtol(s, &s, 10); } }
                    (void)it;
                    if (iv < 0) iv = (long)v.size() + 1 + iv;
                    if (in < 0) in = (long)vn.size() + 1 + in;
                    pos_.push_back(v[iv - 1]);
                    aiVector3D z{0, 0, 0};
                    nrm_.push_back(in > 0 ? vn[in - 1] : z);
                    ++n;
                }
                faces_.push_back(aiFace{(unsigned)n, nullptr});
                (void)first;
            }
        }
        fclose(f);
        idx_.resize(pos_.size());
        for (size_t i = 0; i < idx_.size(); ++i) idx_[i] = (unsigned)i;
        size_t off = 0;
        for (aiFace& fc : faces_) { fc.mIndices = idx_.data() + off; off += fc.mNumIndices; }
        mesh_.mNumVertices = (unsigned)pos_.size(); mesh_.mVertices = pos_.data(); mesh_.mNormals = nrm_.data();
        mesh_.mNumFaces = (unsigned)faces_.size(); mesh_.mFaces = faces_.data();
        meshes_[0] = &mesh_; root_mesh_[0] = 0;
        root_.mNumMeshes = 1; root_.mMeshes = root_mesh_; root_.mNumChildren = 0; root_.mChildren = nullptr;
        scene_.mFlags = 0; scene_.mRootNode = &root_; scene_.mMeshes = meshes_;
        return &scene_;
    }
};
}
