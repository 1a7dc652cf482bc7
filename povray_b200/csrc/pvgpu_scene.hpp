// Internal host-side image of a flattened scene (the tables of include/pvgpu.h) plus the handles of
// their device copies.  Shared by pvgpu_host.cpp (ABI setters, validation, tree build, file I/O) and
// pvgpu_device.cu (upload + kernels).  Nothing here is part of the public ABI.
#pragma once
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>
#include "pvgpu.h"

namespace pvgpu {

struct DeviceScene;   // defined in pvgpu_device.cu

struct Scene {
    pvgpu_globals globals{};
    pvgpu_camera  camera{};
    bool have_camera = false;

    std::vector<pvgpu_object>      objects;
    std::vector<uint32_t>          index_list;
    std::vector<uint32_t>          frame;
    std::vector<pvgpu_transform>   transforms;
    std::vector<pvgpu_node>        nodes;          // scene BBOX_TREE, root = 0 (empty: no tree)

    std::vector<pvgpu_mesh>        meshes;
    std::vector<float>             vertices;       // xyz triples
    std::vector<float>             normals;        // xyz triples
    std::vector<pvgpu_triangle>    triangles;
    std::vector<pvgpu_node>        mesh_nodes;

    std::vector<pvgpu_blob>         blobs;
    std::vector<pvgpu_blob_element> blob_elements;
    std::vector<pvgpu_blob_node>    blob_nodes;
    std::vector<double>             mesh_uv;         // (u, v) pairs (MESH_DATA::UVCoords of all meshes)
    std::vector<uint32_t>           tri_uv;          // per triangle of the triangle table: three indices into mesh_uv (empty: no UV vectors)
    std::vector<int32_t>            blob_textures;   // per blob element: texture index or -1 (empty: no per-component textures)

    std::vector<pvgpu_image>       images;         // image_map pigments (pvgpu_pigment::data = index)
    std::vector<float>             texels;         // r g b filter transmit per texel
    std::vector<double>            shape_data;     // triangle / smooth_triangle / polygon parameters (pvgpu_object::mesh = offset)

    std::vector<pvgpu_light>       lights;
    std::vector<pvgpu_texture>     textures;
    std::vector<pvgpu_pigment>     pigments;
    std::vector<pvgpu_finish>      finishes;
    std::vector<pvgpu_blend_map>   blend_maps;
    std::vector<pvgpu_blend_entry> blend_entries;
    std::vector<pvgpu_warp>        warps;
    std::vector<pvgpu_interior>    interiors;
    std::vector<pvgpu_tnormal>     tnormals;
    std::vector<pvgpu_slope_entry> slope_entries;
    std::vector<pvgpu_sky_sphere>  sky_spheres;    // 0 or 1 entry
    std::vector<pvgpu_fog>         fogs;
    std::vector<float>             irid_wavelengths; // SceneData::iridWavelengths (3 values; empty: no iridescent finish)
    std::vector<double>            camera_ext;     // Camera::Angle, H_Angle, V_Angle (empty: perspective / orthographic only)

    // derived at finalize
    bool     all_shadow_casters_opaque = true;
    int      device = -1;
    size_t   device_bytes = 0;
    DeviceScene* dev = nullptr;               // first device of the scene (owner of device-side outputs); non-null once finalized
    std::vector<DeviceScene*> devs;           // one replica of the tables per device (pvgpu_scene_finalize_multi)
};

// error plumbing (thread-local message)
int  fail(int code, const char* fmt, ...);
void clear_error();

// host helpers implemented in pvgpu_host.cpp
int  validate_scene(Scene& s);

// The reference's tree construction (Build_BBox_Tree, boundingbox.cpp:262-323) on a list of leaf boxes.
// `finite` / `infinite` hold (box, payload) leaves; output nodes are laid out so that the children of a
// node are contiguous and the root is node 0.
struct LeafBox { float lo[3]; float size[3]; uint32_t payload; };
void build_bbox_tree(const std::vector<LeafBox>& finite, const std::vector<LeafBox>& infinite,
                     std::vector<pvgpu_node>& out);

// device side (pvgpu_device.cu)
int  device_upload(Scene& s, const int* devices, int n_devices);     // n_devices <= 0: every visible device
void device_release(Scene& s);

}  // namespace pvgpu
