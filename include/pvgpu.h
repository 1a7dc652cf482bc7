/*
 * pvgpu.h -- C ABI of the B200-native trace path for POV-Ray ("pvgpu").
 *
 * This is the drop-in boundary.  POV-Ray has no plugin/FFI interface for its
 * render path, so the seam is cut where `TraceTask::Run` (source/backend/render/tracetask.cpp:287)
 * would call `TracePixel`/`Trace` (source/core/render/tracepixel.cpp:311, trace.cpp:135): the parsed
 * scene (`SceneData`, source/core/scene/scenedata.h:85-263) is flattened ONCE into the plain tables
 * declared here, and rectangles obtained from `ViewData::GetNextRectangle`
 * (source/backend/scene/view.cpp:236) are rendered by `pvgpu_render`, whose output layout is what
 * `ViewData::CompletedRectangle` (view.cpp:405) expects (row-major RGBT floats).
 *
 * Every table record cites the reference type it is a flat image of.  Numeric types follow the
 * reference: geometry FP64 (DBL), colours / finish / interior FP32 (COLC, SNGL), bounding boxes and
 * mesh vertices FP32 (BBoxScalar, MeshVector).
 *
 * Conventions: plain C, no exceptions across the ABI; every function returns 0 on success or a
 * negative PVGPU_E_* code and leaves a message retrievable with pvgpu_last_error() (thread local).
 * Caller owns every input array (copied during the call); the library owns device memory.
 * There is NO CPU fallback: every render/trace entry point fails with PVGPU_E_NO_DEVICE when no CUDA
 * device is usable.
 */
#ifndef PVGPU_H
#define PVGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVGPU_ABI_VERSION 3
#define PVGPU_FILE_VERSION 1      /* layout of the table records as written by pvgpu_scene_save */

/* ---- error codes ------------------------------------------------------------------------- */
#define PVGPU_OK              0
#define PVGPU_E_INVALID      -1   /* bad argument / inconsistent tables                        */
#define PVGPU_E_UNSUPPORTED  -2   /* scene uses a feature outside the hot-path scope           */
#define PVGPU_E_NO_DEVICE    -3   /* no usable CUDA device (there is no CPU fallback)          */
#define PVGPU_E_CUDA         -4   /* CUDA runtime error                                        */
#define PVGPU_E_IO           -5   /* file could not be read / written                          */
#define PVGPU_E_ABORTED      -6   /* the cooperate callback asked to stop (Task::Cooperate)    */
#define PVGPU_E_OVERFLOW     -7   /* a device-side fixed capacity was exceeded                 */

/* ---- object table ------------------------------------------------------------------------ */

/* Primitive kinds (ObjectBase subclasses, source/core/shape/). */
enum {
    PVGPU_OBJ_SPHERE           = 1,  /* sphere.h:73    p[0..2]=Center p[3]=Radius; aux=Do_Ellipsoid        */
    PVGPU_OBJ_BOX              = 2,  /* box.h:90       p[0..2]=bounds[0] p[3..5]=bounds[1]                  */
    PVGPU_OBJ_PLANE            = 3,  /* plane.h:76     p[0..2]=Normal_Vector p[3]=Distance                  */
    PVGPU_OBJ_QUADRIC          = 4,  /* quadric.h:79   p[0..2]=Square_Terms p[3..5]=Mixed_Terms p[6..8]=Terms p[9]=Constant */
    PVGPU_OBJ_TORUS            = 5,  /* torus.h:80     p[0]=MajorRadius p[1]=MinorRadius; aux=spindle mode (0 = plain torus), p[2]=mSpindleTipYSqr */
    PVGPU_OBJ_MESH             = 6,  /* mesh.h:107     mesh = index into mesh table                         */
    PVGPU_OBJ_CSG_UNION        = 7,  /* csg.h CSGUnion        children in index list                        */
    PVGPU_OBJ_CSG_INTERSECTION = 8,  /* csg.h CSGIntersection (difference = intersection + inverted kids)   */
    PVGPU_OBJ_CSG_MERGE        = 9,  /* csg.h CSGMerge                                                      */
    PVGPU_OBJ_BLOB             = 10, /* blob.h:142     mesh = index into the blob table                     */
    PVGPU_OBJ_CONE             = 11, /* cone.h:66      cone / cylinder in canonical space (transform required); p[0]=dist; CYLINDER / CLOSED flags */
    PVGPU_OBJ_DISC             = 12, /* disc.h:73      p[0..2]=normal p[3]=iradius2 p[4]=oradius2 (transform required)      */
    PVGPU_OBJ_TRIANGLE         = 13, /* triangle.h:73 / :103  mesh = offset into the shape-data table: P1 P2 P3 Normal_Vector Distance
                                        (13 doubles), smooth_triangle: + N1 N2 N3 Perp (25 doubles);
                                        aux = Dominant_Axis | vAxis << 2 | PVGPU_TRIANGLE_SMOOTH                 */
    PVGPU_OBJ_POLYGON          = 14, /* polygon.h:83   p[0..2]=S_Normal; aux = Data->Number; mesh = offset into the shape-data table
                                        (2 doubles per point, Data->Points); transform required                   */
    PVGPU_OBJ_POLY             = 15, /* polynomial.h:78 poly / cubic / quartic of Order <= 4: aux = Order, mesh = offset into the shape-data table
                                        ((Order+1)(Order+2)(Order+3)/6 coefficients, Coeffs); transform required; STURM flag */
    PVGPU_OBJ_GLYPH            = 16, /* truetype.h:95  one character of a text object (the text itself is a CSG union of these): p[0] = depth;
                                        transform required; mesh = offset into the shape-data table: the segment count, then per outline
                                        segment 7 doubles - kind (0 line, 1 quadratic curve), x0 y0, x1 y1, x2 y2 - in the order
                                        GlyphIntersect / Inside_Glyph walk the contours (truetype.cpp:2392-2925), the far end of a curve
                                        whose next point is off-curve already moved to the midpoint, zero-length lines dropped.
                                        Hit aux: bits 0-1 = 0 face z = 0, 1 face z = depth, 2 wall; bit 2 = which root of the wall's
                                        quadratic; bits 3.. = segment */
    PVGPU_OBJ_PRISM            = 17, /* prism.h:98     p[0..1] = Height1 Height2, p[2..5] = x1 y1 x2 y2, p[6..9] = u1 v1 u2 v2 (the spline's
                                        bounding rectangles); aux = Spline_Type (1 linear .. 4 bezier) | Sweep_Type (1 linear, 2 conic) << 4;
                                        transform required; CLOSED / STURM / DEGENERATE flags; mesh = offset into the shape-data table:
                                        Number, then per PRISM_SPLINE_ENTRY 15 doubles: x1 y1 x2 y2, v1 u2 v2, A B C D (x y each).
                                        Hit aux: bits 0-1 = 0 base, 1 cap, 2 spline; bits 2-3 = root index; bits 4.. = segment */
    PVGPU_OBJ_SUPERELLIPSOID   = 18  /* superellipsoid.h:78  p[0..2] = Power (2/e, e/n, 2/n); transform required; aux bit 0 = IS_CHILD_OBJECT
                                        (a CSG operand reports every hit, a stand-alone object stops at the first); no clipped_by */
};
#define PVGPU_OBJ_LAST PVGPU_OBJ_SUPERELLIPSOID
#define PVGPU_TRIANGLE_SMOOTH 0x10u

#define PVGPU_IS_CSG(type) ((type) >= PVGPU_OBJ_CSG_UNION && (type) <= PVGPU_OBJ_CSG_MERGE)

/* Object flags: the reference's ObjectBase::Flags bits verbatim (source/core/scene/object.h:88-117). */
#define PVGPU_NO_SHADOW_FLAG          0x00000001u
#define PVGPU_CLOSED_FLAG             0x00000002u
#define PVGPU_INVERTED_FLAG           0x00000004u
#define PVGPU_CYLINDER_FLAG           0x00000010u
#define PVGPU_DEGENERATE_FLAG         0x00000020u
#define PVGPU_STURM_FLAG              0x00000040u
#define PVGPU_OPAQUE_FLAG             0x00000080u
#define PVGPU_MULTITEXTURE_FLAG       0x00000100u
#define PVGPU_INFINITE_FLAG           0x00000200u
#define PVGPU_HOLLOW_FLAG             0x00000800u
#define PVGPU_UV_FLAG                 0x00002000u
#define PVGPU_DOUBLE_ILLUMINATE_FLAG  0x00004000u
#define PVGPU_NO_IMAGE_FLAG           0x00008000u
#define PVGPU_NO_REFLECTION_FLAG      0x00010000u
#define PVGPU_NO_GLOBAL_LIGHTS_FLAG   0x00020000u
#define PVGPU_CUTAWAY_TEXTURES_FLAG   0x10000000u

/* One ObjectBase (object.h:173-200) plus the subclass parameters. 168 bytes. */
typedef struct pvgpu_object {
    uint32_t type;               /* PVGPU_OBJ_*                                                     */
    uint32_t flags;              /* ObjectBase::Flags                                               */
    int32_t  texture;            /* ObjectBase::Texture          -> texture table index, -1 = none  */
    int32_t  interior_texture;   /* ObjectBase::Interior_Texture -> texture table index, -1 = none  */
    int32_t  interior;           /* ObjectBase::interior         -> interior table index, -1 = none */
    int32_t  transform;          /* ObjectBase::Trans            -> transform table index, -1 = none*/
    int32_t  parent;             /* enclosing CSG object (object index), -1 for frame-level objects */
    uint32_t child_first, child_count;   /* CompoundObject::children -> range in the index list     */
    uint32_t clip_first,  clip_count;    /* ObjectBase::Clip         -> range in the index list     */
    uint32_t bound_first, bound_count;   /* ObjectBase::Bound        -> range in the index list     */
    int32_t  mesh;               /* PVGPU_OBJ_MESH: mesh table index                                */
    uint32_t aux;                /* see PVGPU_OBJ_* comments; blob: 1 = IS_CHILD_OBJECT                */
    float    bbox[6];            /* ObjectBase::BBox: lowerLeft xyz, size xyz (boundingbox.h:93)    */
    uint32_t reserved;
    double   p[10];              /* see PVGPU_OBJ_* comments                                        */
} pvgpu_object;

/* TRANSFORM (source/core/math/matrix.h): row-major 4x4 `matrix` and `inverse`. */
typedef struct pvgpu_transform {
    double matrix[16];
    double inverse[16];
} pvgpu_transform;

/* ---- bounding tree ----------------------------------------------------------------------- */

/* One BBOX_TREE node (boundingbox.h:172-178), children stored contiguously.
 * count > 0: inner node, children are nodes [first, first+count).
 * count == 0: leaf, `first` is the object index (scene tree) or triangle index (mesh tree). */
#define PVGPU_NODE_INFINITE 1u
typedef struct pvgpu_node {
    float    lo[3];              /* BoundingBox::lowerLeft, verbatim FP32 */
    float    size[3];            /* BoundingBox::size,      verbatim FP32 */
    uint32_t first;
    uint16_t count;              /* BBOX_TREE::Entries  */
    uint16_t flags;              /* BBOX_TREE::Infinite */
} pvgpu_node;

/* ---- meshes ------------------------------------------------------------------------------ */

/* MESH_TRIANGLE (source/core/shape/mesh.h:89-106), indices relative to the mesh's own arrays. */
#define PVGPU_TRI_SMOOTH    1u
#define PVGPU_TRI_THREETEX  2u
typedef struct pvgpu_triangle {
    float    perp[3];            /* Perp                                   */
    float    distance;           /* Distance                               */
    int32_t  normal_ind;         /* Normal_Ind                             */
    int32_t  p1, p2, p3;         /* P1..P3                                 */
    int32_t  n1, n2, n3;         /* N1..N3                                 */
    int32_t  texture, texture2, texture3; /* -1 = use the object's texture */
    uint8_t  flags;              /* PVGPU_TRI_*                            */
    uint8_t  dominant_axis;      /* Dominant_Axis                          */
    uint8_t  v_axis;             /* vAxis                                  */
    uint8_t  reserved;
} pvgpu_triangle;

/* MESH_DATA (mesh.h:108-120) + Mesh members (mesh.h:131-145).  Ranges index the scene-wide
 * vertex / normal / triangle / mesh-node / mesh-texture arrays. */
typedef struct pvgpu_mesh {
    uint32_t vertex_first, vertex_count;      /* float[3] each  (MeshVector)  */
    uint32_t normal_first, normal_count;      /* float[3] each                */
    uint32_t triangle_first, triangle_count;
    uint32_t node_first, node_count;          /* mesh BBOX_TREE, root = node_first; count 0 = no tree */
    uint32_t texture_first, texture_count;    /* Mesh::Textures -> range in the index list          */
    uint32_t has_inside_vector;
    uint32_t reserved;
    double   inside_vector[3];
} pvgpu_mesh;

/* ---- blobs ------------------------------------------------------------------------------- */

/* Blob_Element::Type (source/core/shape/blob.h:66-72) */
#define PVGPU_BLOB_SPHERE          2
#define PVGPU_BLOB_CYLINDER        4
#define PVGPU_BLOB_ELLIPSOID       8
#define PVGPU_BLOB_BASE_HEMISPHERE 16
#define PVGPU_BLOB_APEX_HEMISPHERE 32

/* Blob_Element (blob.h:85-100) as Blob::Make_Blob left it (coefficients c[] and transforms already resolved). */
typedef struct pvgpu_blob_element {
    uint32_t type;               /* PVGPU_BLOB_*                                   */
    int32_t  transform;          /* Blob_Element::Trans -> transform table, or -1  */
    double   o[3];               /* O                                              */
    double   len, rad2;          /* len, rad2                                      */
    double   c[3];               /* c[0..2]                                        */
} pvgpu_blob_element;

/* BSPHERE_TREE node (source/core/bounding/boundingsphere.h:66-72), children stored contiguously.
 * count > 0: inner node, children are nodes [first, first+count) of the blob's node range;
 * count == 0: leaf, `first` is the element index (relative to the blob's element range). */
typedef struct pvgpu_blob_node {
    double   c[3];               /* C  */
    double   r2;                 /* r2 */
    uint32_t first, count;
} pvgpu_blob_node;

/* Blob_Data (blob.h:102-118). */
typedef struct pvgpu_blob {
    uint32_t element_first, element_count;
    uint32_t node_first, node_count;         /* node_count == 0: no bounding hierarchy (Blob_Data::Tree == nullptr) */
    double   threshold;                      /* Threshold */
} pvgpu_blob;

/* ---- lights ------------------------------------------------------------------------------ */
enum { PVGPU_LIGHT_POINT = 1, PVGPU_LIGHT_SPOT = 2, PVGPU_LIGHT_FILL = 3, PVGPU_LIGHT_CYLINDER = 4 };
#define PVGPU_LIGHT_AREA             0x001u
#define PVGPU_LIGHT_FULL_AREA        0x002u
#define PVGPU_LIGHT_JITTER           0x004u
#define PVGPU_LIGHT_ORIENT           0x008u
#define PVGPU_LIGHT_CIRCULAR         0x010u
#define PVGPU_LIGHT_PARALLEL         0x020u
#define PVGPU_LIGHT_MEDIA_ATTEN      0x040u
#define PVGPU_LIGHT_MEDIA_INTERACT   0x080u
#define PVGPU_LIGHT_GROUP            0x100u

/* LightSource (source/core/scene/object.h:313-351). */
typedef struct pvgpu_light {
    uint32_t type;               /* Light_Type */
    uint32_t flags;              /* PVGPU_LIGHT_* */
    float    colour[3];
    int32_t  projected_through;  /* object index or -1 */
    double   center[3], direction[3], points_at[3], axis1[3], axis2[3];
    double   coeff, radius, falloff, fade_distance, fade_power;
    int32_t  area_size1, area_size2, adaptive_level;
    uint32_t object_flags;       /* the light's own ObjectBase::Flags */
} pvgpu_light;

/* ---- materials --------------------------------------------------------------------------- */

/* Pattern kinds the device evaluates (source/core/material/pattern.h). */
enum {
    PVGPU_PAT_PLAIN    = 1,      /* PlainPattern / PLAIN_PATTERN: use `colour`            */
    PVGPU_PAT_CHECKER  = 2,      /* CheckerPattern   pattern.cpp:5691                     */
    PVGPU_PAT_BOZO     = 3,      /* NoisePattern/Bozo pattern.cpp:7858                    */
    PVGPU_PAT_GRANITE  = 4,      /* GranitePattern   pattern.cpp:6429                     */
    PVGPU_PAT_GRADIENT = 5,      /* GradientPattern  pattern.cpp:6386, p[0..2]=gradient   */
    PVGPU_PAT_MARBLE   = 6,      /* MarblePattern    pattern.cpp:7831                     */
    PVGPU_PAT_WRINKLES = 7,      /* WrinklesPattern  pattern.cpp:8720                     */
    PVGPU_PAT_ONION    = 8,      /* OnionPattern     pattern.cpp:7934                     */
    PVGPU_PAT_BRICK    = 9,      /* BrickPattern     pattern.cpp:5495, p[0..2]=brick size p[3]=mortar */
    PVGPU_PAT_HEXAGON  = 10,     /* HexagonPattern   pattern.cpp:6512                     */
    PVGPU_PAT_SPOTTED  = 11,     /* SpottedPattern (== bozo noise)                        */
    PVGPU_PAT_AGATE    = 12,     /* AgatePattern     pattern.cpp:5396, p[0]=agateTurbScale */
    PVGPU_PAT_WOOD     = 13,     /* WoodPattern      pattern.cpp:8651                     */
    PVGPU_PAT_LEOPARD  = 14,     /* LeopardPattern   pattern.cpp:7179                     */
    PVGPU_PAT_SPHERICAL= 15,     /* SphericalPattern pattern.cpp:8551                     */
    PVGPU_PAT_BOXED    = 16,     /* BoxedPattern     pattern.cpp:5454                     */
    PVGPU_PAT_RADIAL   = 17,     /* RadialPattern    pattern.cpp:8115                     */
    PVGPU_PAT_CYLINDRICAL = 18,  /* CylindricalPattern pattern.cpp:6025                   */
    PVGPU_PAT_PLANAR   = 19,     /* PlanarPattern    pattern.cpp:8024                     */
    PVGPU_PAT_DENTS    = 20,     /* DentsPattern     pattern.cpp:6307                     */
    PVGPU_PAT_RIPPLES  = 21,     /* RipplesPattern   pattern.cpp:8163                     */
    PVGPU_PAT_WAVES    = 22,     /* WavesPattern     pattern.cpp:8593                     */
    PVGPU_PAT_QUILTED  = 23,     /* QuiltedPattern   pattern.cpp:8067, p[0..1] = Control0, Control1 */
    PVGPU_PAT_AVERAGE  = 24,     /* AVERAGE_PATTERN pigment: weighted mean of the blend map's entries (pigment.cpp:566-596) */
    PVGPU_PAT_CRACKLE  = 25,     /* CracklePattern   pattern.cpp:5760; `data` = offset into the shape-data table of 9 doubles:
                                    crackleForm xyz, crackleMetric, crackleOffset, crackleIsSolid, repeat xyz */
    PVGPU_PAT_CELLS    = 26,     /* CellsPattern     pattern.cpp:5652 */
    PVGPU_PAT_IMAGE_MAP = 27,    /* IMAGE_MAP_PATTERN pigment (ColourImagePattern, pattern.cpp:493-514); `data` = index into the image table */
    PVGPU_PAT_FRACTAL  = 28,     /* FractalPattern family (pattern.cpp:6895-7098, 7228-7751): `data` = offset into the shape-data table of
                                    8 doubles: PVGPU_FRACTAL_* kind, maxIterations, exteriorType, interiorType, exteriorFactor,
                                    interiorFactor, juliaCoord u v */
    PVGPU_PAT_SPIRAL1  = 29,     /* Spiral1Pattern   pattern.cpp:8396, p[0] = arms */
    PVGPU_PAT_SPIRAL2  = 30,     /* Spiral2Pattern   pattern.cpp:8473, p[0] = arms */
    PVGPU_PAT_PIGMENT  = 32,     /* PigmentPattern (`pigment_pattern { ... }`, pattern.cpp:7974-7990): `data` = index of the pigment whose
                                    greyscale is the pattern value; as the pattern of a pigment or of a normal */
    PVGPU_PAT_UV_MAP   = 31      /* UV_MAP_PATTERN pigment (`pigment { uv_mapping ... }`, PigmentBlendMap::ComputeUVMapped, pigment.cpp:603-618):
                                    `data` = index of the pigment that is evaluated at (u, v, 0) of the hit */
};
#define PVGPU_PAT_LAST PVGPU_PAT_PIGMENT
/* iteration formulas of PVGPU_PAT_FRACTAL (exponents above 4 - MandelXPattern / JuliaXPattern - are not served) */
enum { PVGPU_FRACTAL_MANDEL2 = 0, PVGPU_FRACTAL_MANDEL3 = 1, PVGPU_FRACTAL_MANDEL4 = 2, PVGPU_FRACTAL_JULIA2 = 3, PVGPU_FRACTAL_JULIA3 = 4,
       PVGPU_FRACTAL_JULIA4 = 5, PVGPU_FRACTAL_MAGNET1M = 6, PVGPU_FRACTAL_MAGNET1J = 7, PVGPU_FRACTAL_MAGNET2M = 8, PVGPU_FRACTAL_MAGNET2J = 9 };
/* ContinuousPattern::waveType (pattern.h:108-117) */
enum { PVGPU_WAVE_RAW = 0, PVGPU_WAVE_RAMP = 1, PVGPU_WAVE_SINE = 2, PVGPU_WAVE_TRIANGLE = 3,
       PVGPU_WAVE_SCALLOP = 4, PVGPU_WAVE_CUBIC = 5, PVGPU_WAVE_POLY = 6 };

/* Warps (source/core/material/warp.h). */
enum { PVGPU_WARP_TRANSFORM = 1, PVGPU_WARP_TURBULENCE = 2, PVGPU_WARP_CLASSIC_TURBULENCE = 3,
       /* the point-mapping warps (warp.cpp:124-545): parameters as doubles in the shape-data table at offset `transform` */
       PVGPU_WARP_BLACK_HOLE = 4,   /* Center xyz, Repeat_Vector xyz, Strength, Radius, Power, flags (1 Inverted, 2 Repeat), Type: 11 doubles;
                                       `uncertain` black holes (WarpRands) are not served */
       PVGPU_WARP_REPEAT = 5,       /* Axis, Width, Flip xyz, Offset xyz: 8 doubles */
       PVGPU_WARP_CUBIC = 6,        /* no parameters */
       PVGPU_WARP_CYLINDRICAL = 7,  /* Orientation_Vector xyz, DistExp: 4 doubles */
       PVGPU_WARP_SPHERICAL = 8,    /* Orientation_Vector xyz, DistExp: 4 doubles */
       PVGPU_WARP_TOROIDAL = 9,     /* Orientation_Vector xyz, DistExp, MajorRadius: 5 doubles */
       PVGPU_WARP_PLANAR = 10 };    /* Orientation_Vector xyz, OffSet: 4 doubles */
typedef struct pvgpu_warp {
    uint32_t type;
    int32_t  transform;          /* TransformWarp::Trans -> transform table index; warps >= PVGPU_WARP_BLACK_HOLE: shape-data offset */
    double   turbulence[3];      /* GenericTurbulenceWarp::Turbulence             */
    int32_t  octaves;
    float    lambda, omega;
    uint32_t handled_by_pattern; /* ClassicTurbulence::handledByPattern           */
} pvgpu_warp;

/* BlendMapEntry<TransColour> (pigment.h:125): value + rgb,filter,transmit. */
typedef struct pvgpu_blend_entry {
    float value;
    float colour[5];
} pvgpu_blend_entry;

/* ColourBlendMap / PigmentBlendMap: contiguous entry range + GenericPigmentBlendMap::blendMode/blendGamma.
 * blend_mode & PVGPU_BLEND_PIGMENT_MAP: a pigment_map (BlendMapEntry<PIGMENT*>, pigment.h:113-123) - every entry's colour[0]
 * holds the pigment table index of its PIGMENT (exact in FP32, indices < 2^24), evaluated at the parent's warped point. */
#define PVGPU_BLEND_PIGMENT_MAP 0x100
/* blend_mode & PVGPU_BLEND_TEXTURE_MAP: a texture_map (BlendMapEntry<TexturePtr>, texture.h:96-104): colour[0] = texture table index */
#define PVGPU_BLEND_TEXTURE_MAP 0x200
/* blend_mode & PVGPU_BLEND_NORMAL_MAP: a normal_map (BlendMapEntry<TNORMAL*>, normal.h:98-110): colour[0] = tnormal table index */
#define PVGPU_BLEND_NORMAL_MAP  0x400
typedef struct pvgpu_blend_map {
    uint32_t entry_first, entry_count;
    int32_t  blend_mode;
    float    blend_gamma;
} pvgpu_blend_map;

/* PIGMENT (pigment.h:130-135) + the pattern object it owns (pattern.h BasicPattern/ContinuousPattern). */
typedef struct pvgpu_pigment {
    uint32_t pattern;            /* PVGPU_PAT_*                                     */
    uint32_t wave_type;          /* PVGPU_WAVE_*                                    */
    float    frequency, phase, exponent;     /* waveFrequency, wavePhase, waveExponent */
    int32_t  noise_generator;    /* BasicPattern::noiseGenerator, 0 = scene default */
    uint32_t warp_first, warp_count;         /* BasicPattern::warps -> warp table range */
    int32_t  blend_map;          /* Blend_Map -> blend map index, -1 = none         */
    float    colour[5];          /* PIGMENT::colour                                 */
    float    quick_colour[5];    /* PIGMENT::Quick_Colour (NaN red = invalid)       */
    uint32_t data;               /* pattern parameters that do not fit p[]: offset into the shape-data table (crackle) */
    double   p[4];               /* pattern-specific, see PVGPU_PAT_*               */
} pvgpu_pigment;

/* ImageData (source/core/support/imageutil.h:104-131) of an image_map pigment.  The texels are handed over decoded: one
 * r g b filter transmit float quintuple per pixel as Image::GetRGBFTValue returns it (file gamma, palette and per-index
 * filter / transmit applied by the reference's image readers), row-major, row 0 = top row of the file. */
#define PVGPU_IMAGE_ONCE          1u   /* Once_Flag                                                                            */
#define PVGPU_IMAGE_PREMULTIPLIED 2u   /* texels are alpha-premultiplied (Image::IsPremultiplied): un-premultiplied after interpolation */
#define PVGPU_IMAGE_TRANSMIT_ALL  4u   /* proper "transmit all / filter all" (imageutil.cpp:400-402): scaled by the image's alpha;
                                          the legacy mode (added to every texel) is applied to the texels by the caller       */
typedef struct pvgpu_image {
    uint32_t width, height;      /* iwidth, iheight                                 */
    uint32_t map_type;           /* Map_Type: 0 planar, 1 spherical, 2 cylindrical, 5 torus, 7 angular (imageutil.h:75-82) */
    uint32_t interpolation;      /* Interpolation_Type: 0 none, 2 bilinear, 3 bicubic, 4 normalized distance (imageutil.h:89-93) */
    uint32_t flags;              /* PVGPU_IMAGE_*                                   */
    uint32_t data_first;         /* first texel of this image in the texel table    */
    float    fwidth, fheight;    /* width, height (SNGL)                            */
    float    all_filter, all_transmit;   /* AllFilter, AllTransmit                  */
    double   gradient[3];        /* Gradient                                        */
    double   offset[2];          /* Offset                                          */
} pvgpu_image;

/* FINISH (source/core/material/texture.h:119-142), same field order. */
typedef struct pvgpu_finish {
    float diffuse, diffuse_back, brilliance, brilliance_adjust, brilliance_adjust_rad;
    float specular, roughness, phong, phong_size;
    float irid, irid_film_thickness, irid_turb;
    float reflect_exp, crand, metallic;
    float ambient[3], emission[3], reflection_max[3], reflection_min[3];
    float reflection_falloff;
    float fresnel;
    float reflect_metallic;
    int32_t reflection_fresnel;
    int32_t conserve_energy;
    int32_t alpha_knockout;
    int32_t use_subsurface;
} pvgpu_finish;

/* ---- normal perturbation (source/core/material/normal.cpp, normal.h:118-123) ---------------- */
enum {
    PVGPU_NORM_BUMPS    = 1,     /* bumps    normal.cpp:235 */
    PVGPU_NORM_DENTS    = 2,     /* dents    normal.cpp:272 */
    PVGPU_NORM_RIPPLES  = 3,     /* ripples  normal.cpp:130 (wave sources: Initialize_Waves, noise.cpp:189) */
    PVGPU_NORM_WAVES    = 4,     /* waves    normal.cpp:180 */
    PVGPU_NORM_WRINKLES = 5,     /* wrinkles normal.cpp:325 */
    PVGPU_NORM_QUILTED  = 6,     /* quilted  normal.cpp:371, carrier p[0..1] = Control0, Control1 */
    PVGPU_NORM_PATTERN  = 7,     /* any continuous pattern: 4 samples on Pyramid_Vect + slope map (normal.cpp:893-918); with `normal_map`
                                    set: any pattern (block patterns too) selecting / blending whole normals (normal.cpp:824-848) */
    PVGPU_NORM_AVERAGE  = 8      /* average normal_map (NormalBlendMap::ComputeAverage, normal.cpp:1033-1059) */
};
#define PVGPU_DONT_SCALE_BUMPS_FLAG 8u   /* pattern.h:106 */

/* BlendMapEntry<Vector2d> of a SlopeBlendMap (normal.h:84-96). */
typedef struct pvgpu_slope_entry {
    float    value;
    uint32_t reserved;
    double   height, slope;      /* Vals[0], Vals[1] */
} pvgpu_slope_entry;

/* TNORMAL (normal.h:118-123).  The TPATTERN base it shares with PIGMENT (pattern kind, waveform, frequency / phase, warps,
 * noise generator, pattern parameters) is stored as a pvgpu_pigment record used as pattern carrier (its colours are unused). */
typedef struct pvgpu_tnormal {
    uint32_t type;               /* PVGPU_NORM_*                                         */
    uint32_t flags;              /* PVGPU_DONT_SCALE_BUMPS_FLAG                          */
    int32_t  pattern;            /* pigment table index of the pattern carrier           */
    uint32_t slope_first, slope_count;   /* slope_map entries, count 0 = none            */
    float    amount, delta;      /* Amount, Delta                                        */
    uint32_t normal_map;         /* 0 = none, else blend map index + 1 of a normal_map (PVGPU_BLEND_NORMAL_MAP) */
} pvgpu_tnormal;

/* TEXTURE (texture.h:108-117): one layer; `next` chains the layers of a layered texture. */
typedef struct pvgpu_texture {
    uint32_t type;               /* PVGPU_PAT_PLAIN: one layer (pigment / finish / tnormal).  Any other PVGPU_PAT_*: a patterned texture
                                    (texture_map; PVGPU_PAT_AVERAGE: average texture_map) - `pigment` then is the pattern carrier
                                    (a pvgpu_pigment record holding the TPATTERN part) and `blend_map` the texture map        */
    int32_t  next;               /* TEXTURE::Next, -1 = last layer                       */
    int32_t  pigment;
    int32_t  finish;
    int32_t  tnormal;            /* TEXTURE::Tnormal -> tnormal table index, -1 = none     */
    int32_t  blend_map;          /* patterned textures: blend map index (PVGPU_BLEND_TEXTURE_MAP); plain layers: unused (0) */
} pvgpu_texture;

/* Interior (source/core/coretypes.h:193-214). */
typedef struct pvgpu_interior {
    int32_t hollow;
    int32_t disp_nelems;
    float   ior, dispersion, caustics, old_refract, fade_distance, fade_power;
    float   fade_colour[3];
    uint32_t reserved;
} pvgpu_interior;

/* ---- atmosphere: sky_sphere and fog (source/core/scene/atmosphere.h) ---------------------- */

/* Skysphere_Struct (atmosphere.h:110-117). */
typedef struct pvgpu_sky_sphere {
    uint32_t pigment_first, pigment_count;   /* Pigments -> range in the index list (entries = pigment table indices) */
    int32_t  transform;                      /* Trans -> transform table index, -1 = none                            */
    float    emission[3];                    /* Emission                                                             */
} pvgpu_sky_sphere;

enum { PVGPU_FOG_CONSTANT = 1, PVGPU_FOG_GROUND = 2 };   /* ORIG_FOG, GROUND_MIST (atmosphere.h:71-72) */
/* Fog_Struct (atmosphere.h:81-94); the list SceneData::fog is passed in list order. */
typedef struct pvgpu_fog {
    uint32_t type;               /* PVGPU_FOG_*                                                   */
    int32_t  turbulence;         /* Turb -> warp table index of a turbulence warp, -1 = none      */
    double   distance, alt, offset;
    double   up[3];
    float    colour[5];          /* rgb, filter, transmit                                         */
    float    turb_depth;
} pvgpu_fog;

/* ---- frame-level ------------------------------------------------------------------------- */

/* QualityFlags (source/core/coretypes.h:558-585) */
#define PVGPU_Q_AMBIENT_ONLY 0x001u
#define PVGPU_Q_QUICK_COLOUR 0x002u
#define PVGPU_Q_SHADOWS      0x004u
#define PVGPU_Q_AREA_LIGHTS  0x008u
#define PVGPU_Q_REFRACTIONS  0x010u
#define PVGPU_Q_REFLECTIONS  0x020u
#define PVGPU_Q_NORMALS      0x040u
#define PVGPU_Q_MEDIA        0x080u   /* gates fog (trace.cpp:207-216); participating media itself is out of scope */
#define PVGPU_Q_DEFAULT (PVGPU_Q_SHADOWS | PVGPU_Q_AREA_LIGHTS | PVGPU_Q_REFRACTIONS | PVGPU_Q_REFLECTIONS | PVGPU_Q_NORMALS | PVGPU_Q_MEDIA)

/* The SceneData scalars the trace path reads (scenedata.h:85-263). */
typedef struct pvgpu_globals {
    uint32_t max_trace_level;    /* parsedMaxTraceLevel  */
    uint32_t language_version;   /* EffectiveLanguageVersion(), e.g. 370 / 380 */
    int32_t  noise_generator;    /* noiseGenerator       */
    uint32_t bounding_method;    /* boundingMethod: 0 = linear object loop, 1 = BBOX_TREE */
    uint32_t quality_flags;      /* PVGPU_Q_*            */
    int32_t  output_alpha;       /* outputAlpha          */
    double   adc_bailout;        /* parsedAdcBailout     */
    float    ambient_light[3];   /* ambientLight         */
    float    background[5];      /* backgroundColour (rgb, filter, transmit) */
    float    atmosphere_ior, atmosphere_dispersion;
    uint32_t number_of_waves;
    uint32_t reserved;
} pvgpu_globals;

/* Camera (source/core/scene/camera.h:84-136), the members TracePixel reads (tracepixel.cpp:235-391). */
/* camera.h:71-81; types 3..11 additionally read Camera::Angle / H_Angle / V_Angle (pvgpu_scene_set_camera_angles) */
enum { PVGPU_CAMERA_PERSPECTIVE = 1, PVGPU_CAMERA_ORTHOGRAPHIC = 2, PVGPU_CAMERA_FISHEYE = 3, PVGPU_CAMERA_ULTRA_WIDE_ANGLE = 4,
       PVGPU_CAMERA_OMNIMAX = 5, PVGPU_CAMERA_PANORAMIC = 6, PVGPU_CAMERA_CYL_1 = 7, PVGPU_CAMERA_CYL_2 = 8, PVGPU_CAMERA_CYL_3 = 9,
       PVGPU_CAMERA_CYL_4 = 10, PVGPU_CAMERA_SPHERICAL = 11 };
typedef struct pvgpu_camera {
    uint32_t type;
    uint32_t reserved;           /* camera { normal { ... } } (Camera::Tnormal, tracepixel.cpp:917-924): index into the tnormal table + 1,
                                    0 = none; perspective and orthographic cameras only */
    double   location[3], direction[3], up[3], right[3];
    double   max_ray_distance;
} pvgpu_camera;

/* TraceTask constructor arguments (source/backend/render/tracetask.h:73-75). */
typedef struct pvgpu_aa {
    uint32_t method;             /* tracingMethod: 0 none, 1 non-adaptive (+AM1), 2 adaptive (+AM2) */
    uint32_t depth;              /* aaDepth      (+R)  */
    double   threshold;          /* aaThreshold  (+A)  */
    double   jitter_scale;       /* jitterScale  (+J), 0 = off */
    double   gamma;              /* aaGamma      (+AG), default 2.5 (view.cpp:724) */
} pvgpu_aa;

/* POVRect (inclusive corners, source/backend/frame.h). */
typedef struct pvgpu_rect { int32_t left, top, right, bottom; } pvgpu_rect;

/* Device counters, same meaning as the reference's RenderStatistics ids (statisticids.h:120-253). */
typedef struct pvgpu_stats {
    uint64_t rays;               /* Number_Of_Rays   (every TraceRay call)        */
    uint64_t shadow_ray_tests;   /* Shadow_Ray_Tests (every shadow traversal)     */
    uint64_t reflected_rays, refracted_rays, transmitted_rays, tir_rays;
    uint64_t adc_saves;
    uint64_t samples;            /* AA samples beyond one per pixel               */
    uint64_t waves;              /* wavefront iterations                          */
    uint64_t kernel_launches;    /* CUDA kernels launched by the call             */
    uint32_t max_trace_level;    /* highest level reached                         */
    uint32_t overflow;           /* non-zero: a device capacity was exceeded      */
    /* work counters of the traversal kernels: bounding-box slab tests (scene tree + mesh trees) and primitive tests (top-level
     * All_Intersections calls + mesh triangles), separately for k_closest and k_shadow_*; bench.py's roofline uses them */
    uint64_t node_tests_closest, prim_tests_closest, node_tests_shadow, prim_tests_shadow;
    double   device_ms;          /* CUDA-event time of the device work            */
    /* per kernel family, timed with CUDA events on the launching stream:
     * 0 camera rays (k_primary), 1 closest hit (k_closest), 2 shading (k_shade), 3 shadow rays (k_shadow_*),
     * 4 anti-aliasing bookkeeping.  kernel_items = rays (samples) the launches processed. */
    double   kernel_ms[5];
    uint64_t kernel_count[5];
    uint64_t kernel_items[5];
} pvgpu_stats;

typedef struct pvgpu_scene pvgpu_scene;   /* opaque */

/* ---- scene construction (replaces: Scene::StartParser handing SceneData to the views) ----- */
int  pvgpu_abi_version(void);
const char* pvgpu_last_error(void);

int  pvgpu_scene_create(pvgpu_scene** out, const pvgpu_globals* g);
void pvgpu_scene_destroy(pvgpu_scene* s);

/* SceneData::objects (all objects incl. CSG children / clip / bound objects; frame-level ones are
 * those listed in `frame`, in SceneData::objects order) + the shared index list. */
int  pvgpu_scene_set_objects(pvgpu_scene* s, const pvgpu_object* objs, size_t n_objs,
                             const uint32_t* index_list, size_t n_index,
                             const uint32_t* frame, size_t n_frame);
int  pvgpu_scene_set_transforms(pvgpu_scene* s, const pvgpu_transform* t, size_t n);
/* SceneData::boundingSlabs (BBOX_TREE built by BoundingTask, boundingtask.cpp:168-209); root = node 0. */
int  pvgpu_scene_set_tree(pvgpu_scene* s, const pvgpu_node* nodes, size_t n);
/* Builds the tree with the reference's own algorithm (Build_Bounding_Slabs, boundingbox.cpp:325-430)
 * from the frame-level objects' bboxes; for callers that do not come from the POV-Ray parser. */
int  pvgpu_scene_build_tree(pvgpu_scene* s);
int  pvgpu_scene_set_meshes(pvgpu_scene* s, const pvgpu_mesh* meshes, size_t n_meshes,
                            const float* vertices, size_t n_vertices,
                            const float* normals, size_t n_normals,
                            const pvgpu_triangle* tris, size_t n_tris,
                            const pvgpu_node* nodes, size_t n_nodes);
/* Blobs: element c[] / transforms as computed by Blob::Make_Blob, bounding-sphere trees as built by
 * Blob::build_bounding_hierarchy (blob.cpp:2516-2766). */
int  pvgpu_scene_set_blobs(pvgpu_scene* s, const pvgpu_blob* blobs, size_t n_blobs,
                           const pvgpu_blob_element* elements, size_t n_elements,
                           const pvgpu_blob_node* nodes, size_t n_nodes);
/* Per-component textures of blobs (Blob::Element_Texture, blob.h:167): one texture index per blob element, -1 = the object's own
 * texture.  Objects whose blob has any carry MULTITEXTURE_FLAG; Blob::Determine_Textures (blob.cpp:2768-2843) then blends the
 * components' textures by their field contribution at the hit point.  n must equal the number of blob elements (or 0). */
int  pvgpu_scene_set_blob_textures(pvgpu_scene* s, const int32_t* textures, size_t n);
/* UV vectors of meshes (MESH_DATA::UVCoords, MESH_TRIANGLE::UV1..UV3; Mesh::UVCoord, mesh.cpp:2256-2332): `uv` holds n_uv (u, v) pairs,
   `tri_uv` three indices into it for every triangle of the triangle table (n_tri = 0: no mesh has UV vectors).  Used by objects with
   PVGPU_UV_FLAG (`uv_mapping` in an object's texture: ObjectBase::UVCoord replaces the intersection point for every texture
   evaluation, trace.cpp:500-512) and by PVGPU_PAT_UV_MAP pigments.  UVCoord is served for spheres, boxes, tori and meshes; every
   other primitive takes ObjectBase::UVCoord (x, y of the point) except cones / cylinders, which are rejected with the flag. */
int  pvgpu_scene_set_mesh_uv(pvgpu_scene* s, const double* uv, size_t n_uv, const uint32_t* tri_uv, size_t n_tri);
/* image_map pigments: the image table and the texel table (5 floats per texel) its records point into. */
int  pvgpu_scene_set_images(pvgpu_scene* s, const pvgpu_image* images, size_t n_images, const float* texels, size_t n_texel_floats);
/* Shape-data table: FP64 parameters of the primitives whose record does not fit pvgpu_object::p (triangle, smooth_triangle,
 * polygon); an object's `mesh` field is its offset into this array. */
int  pvgpu_scene_set_shape_data(pvgpu_scene* s, const double* data, size_t n);
int  pvgpu_scene_set_lights(pvgpu_scene* s, const pvgpu_light* l, size_t n);
int  pvgpu_scene_set_materials(pvgpu_scene* s,
                               const pvgpu_texture* tex, size_t n_tex,
                               const pvgpu_pigment* pig, size_t n_pig,
                               const pvgpu_finish* fin, size_t n_fin,
                               const pvgpu_blend_map* maps, size_t n_maps,
                               const pvgpu_blend_entry* entries, size_t n_entries,
                               const pvgpu_warp* warps, size_t n_warps,
                               const pvgpu_interior* interiors, size_t n_interiors);
/* Normal perturbations referenced by pvgpu_texture::tnormal, and the slope_map entries they use. */
int  pvgpu_scene_set_normals(pvgpu_scene* s, const pvgpu_tnormal* tn, size_t n_tn, const pvgpu_slope_entry* slopes, size_t n_slopes);
/* SceneData::iridWavelengths (scenedata.h:113; global_settings irid_wavelength) for finishes with iridescence. */
int  pvgpu_scene_set_irid_wavelengths(pvgpu_scene* s, const float wavelengths[3]);
/* SceneData::skysphere (NULL = none) and SceneData::fog (list order). */
int  pvgpu_scene_set_atmosphere(pvgpu_scene* s, const pvgpu_sky_sphere* sky, const pvgpu_fog* fogs, size_t n_fogs);
int  pvgpu_scene_set_camera(pvgpu_scene* s, const pvgpu_camera* cam);
int  pvgpu_scene_get_camera(const pvgpu_scene* s, pvgpu_camera* cam);
/* Camera::Angle, H_Angle, V_Angle (camera.h:105-107, degrees) for the fisheye / ultra_wide_angle / cylinder / spherical cameras. */
int  pvgpu_scene_set_camera_angles(pvgpu_scene* s, double angle, double h_angle, double v_angle);

/* Host-side mesh helper mirroring the parser's mesh2 post-processing (Mesh::Compute_Mesh_Triangle,
 * mesh.cpp:838-954, and Mesh::Build_Mesh_BBox_Tree, mesh.cpp:1376-1413): fills triangle records
 * (normal, Distance, Dominant_Axis, vertex order) and builds the mesh tree.  `indices` = 3 per face.
 * On return *out_mesh is the index of the new mesh in the scene's mesh table. */
int  pvgpu_scene_add_mesh2(pvgpu_scene* s, const double* vertices, size_t n_vertices,
                           const int32_t* indices, size_t n_faces, int32_t* out_mesh);

/* Validates, derives device layouts and uploads everything to CUDA device `device`. */
int  pvgpu_scene_finalize(pvgpu_scene* s, int device);
/* Same, but the scene tables are replicated on `n_devices` CUDA devices of this node (devices[] lists them, NULL = 0 .. n_devices - 1;
 * n_devices <= 0 = every visible device).  pvgpu_render / pvgpu_render_device then shard the rectangles of a call over the
 * devices: one host thread per device takes chunks of rectangles from one atomic counter - the GPU counterpart of the render
 * threads of View::StartRender pulling from ViewData::GetNextRectangle (view.cpp:236-271, 1186-1190) - and finished tiles are
 * gathered into the caller's frame (D2H, or peer copies into devices[0]'s memory for pvgpu_render_device).  Pixels do not depend
 * on the number of devices. */
int  pvgpu_scene_finalize_multi(pvgpu_scene* s, const int* devices, int n_devices);
/* Number of devices the scene lives on (0 before finalize). */
int  pvgpu_scene_device_count(const pvgpu_scene* s);
/* Total bytes uploaded by finalize (the H2D traffic of one scene). */
size_t pvgpu_scene_device_bytes(const pvgpu_scene* s);

/* Flat-scene file (little endian dump of the tables; used by the reference-side adapter and tests). */
int  pvgpu_scene_save(const pvgpu_scene* s, const char* path);
int  pvgpu_scene_load(pvgpu_scene** out, const char* path);

/* ---- rendering (replaces TraceTask::Run: tracetask.cpp:287-657) --------------------------- */

/* Renders `n_rects` rectangles of a width x height image.  rgbt_out (HOST memory) receives, rectangle
 * after rectangle, rect-area pixels row-major, 4 floats each (r, g, b, transm) exactly as
 * ViewData::CompletedRectangle takes them.  `cooperate` (may be NULL) is polled between wavefront
 * iterations like Task::Cooperate(); a non-zero return aborts with PVGPU_E_ABORTED. */
int  pvgpu_render(pvgpu_scene* s, const pvgpu_aa* aa, int width, int height,
                  const pvgpu_rect* rects, size_t n_rects, float* rgbt_out,
                  pvgpu_stats* stats, int (*cooperate)(void*), void* user);

/* Page-locked host memory for rgbt_out: pvgpu_render copies the frame straight into such a buffer with one DMA
 * transfer; any other host pointer is served through an internal pinned staging buffer and an extra memcpy. */
void* pvgpu_host_alloc(size_t bytes);
void  pvgpu_host_free(void* p);

/* Same, but the result stays in DEVICE memory (`d_rgbt_out` is a device pointer with room for the
 * rect-area sum x 4 floats) and the work is enqueued on `cuda_stream` (a cudaStream_t, 0 = default). */
int  pvgpu_render_device(pvgpu_scene* s, const pvgpu_aa* aa, int width, int height,
                         const pvgpu_rect* rects, size_t n_rects, float* d_rgbt_out,
                         pvgpu_stats* stats, void* cuda_stream);

/* Ray-level harness (mirrors Trace::FindIntersection(Intersection&, const Ray&), trace.h:255, with the
 * primary-ray conditions of TraceRay): org_dir = 6 doubles per ray (origin, direction; host memory).
 * obj[i] = frame-or-child object index of Intersection::Object or 0xFFFFFFFF, depth[i] = Depth,
 * aux[i] = Intersection::i1 (box side) / triangle index (mesh), may be NULL. */
int  pvgpu_trace_rays(pvgpu_scene* s, const double* org_dir, size_t n,
                      uint32_t* obj, double* depth, uint32_t* aux);

/* Component-level harnesses (device code of the trace path run on explicit inputs; host arrays):
 *  pvgpu_solve_polynomial mirrors Solve_Polynomial(n, c, r, sturm, epsilon) (source/core/math/polynomialsolver.h:73) for
 *    n_polys polynomials of degree <= 4: coeffs = 5 doubles each (c[4 - degree ..] used, highest power first), degree / sturm per
 *    polynomial; roots = 4 doubles each, counts = number of roots.
 *  pvgpu_noise mirrors Noise / DNoise / Turbulence (source/core/material/noise.h:196-202) at n points: xyz = 3 doubles each,
 *    generator and octaves per point (lambda 2, omega 0.5); out = 5 doubles each: noise, dnoise xyz, turbulence. */
int  pvgpu_solve_polynomial(pvgpu_scene* s, size_t n_polys, const int32_t* degree, const int32_t* sturm, const double* epsilon,
                            const double* coeffs, double* roots, int32_t* counts);
int  pvgpu_noise(pvgpu_scene* s, size_t n, const double* xyz, const int32_t* generator, const int32_t* octaves, double* out);

/* Camera rays exactly as TracePixel::CreateCameraRay makes them for pixel-space (x, y): 6 doubles each. */
int  pvgpu_camera_rays(pvgpu_scene* s, int width, int height, const double* xy, size_t n, double* org_dir);

/* Starts the CUDA driver / context initialisation of `device` on a background thread and returns at once (a one-shot program
 * calls it first thing, so that the ~1 s of cuInit + context creation overlaps the parser instead of preceding the first frame);
 * pvgpu_scene_finalize waits for it.  Optional: finalize initialises on demand. */
void pvgpu_prewarm(int device);

/* Measured FP64 vector peak of CUDA device `device` in TFLOP/s (independent DFMA chains, timed with CUDA events): the second
 * roofline denominator of the trace path.  No scene needed. */
int  pvgpu_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* PVGPU_H */
