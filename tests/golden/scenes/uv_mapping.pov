// uv_mapping: ObjectBase::UVCoord replaces the intersection point in texture evaluation (trace.cpp:500-512, 2351-2362) - at object
// level (the whole texture) and at pigment level (pigment.cpp:603-618); Sphere / Box / Torus / Mesh::UVCoord and the default (x, y)
#version 3.7;
global_settings { assumed_gamma 1.0 max_trace_level 5 }
camera { location <0, 4, -9.5> look_at <0, 0.9, 0> angle 44 }
light_source { <-7, 10, -8> rgb <1, 0.95, 0.9> }
light_source { <8, 5, -5> rgb <0.3, 0.35, 0.5> }
plane { y, 0 pigment { rgb <0.8, 0.8, 0.75> } }
#declare Chk = pigment { checker rgb <0.9, 0.2, 0.2>, rgb <0.95, 0.95, 0.6> scale <0.1, 0.1, 1> }
#declare Grad = pigment { gradient x color_map { [0 rgb <0.1, 0.2, 0.8>] [0.5 rgb <0.9, 0.9, 0.2>] [1 rgb <0.1, 0.7, 0.3>] } frequency 3 }
sphere { 0, 1 uv_mapping texture { pigment { Chk } finish { phong 0.5 } } rotate <20, 40, 0> translate <-3.6, 1, 1.5> }
sphere { 0, 1 pigment { uv_mapping Grad } scale <1, 0.7, 0.8> rotate 30 * z translate <-1.3, 0.8, 2> }
box { -1, 1 uv_mapping texture { pigment { Chk scale 0.7 } } rotate <0, 30, 0> scale 0.75 translate <1.1, 0.75, 1.8> }
torus { 0.8, 0.3 pigment { uv_mapping Grad } rotate <-50, 20, 0> translate <3.5, 1.1, 1.5> }
quadric { <1, 0, 1>, <0, 0, 0>, <0, 0, 0>, -0.25 clipped_by { box { <-1, 0, -1>, <1, 1.5, 1> } } pigment { uv_mapping Chk scale <3, 3, 1> } translate <-3.2, 0, -1.5> }   // the default UVCoord: (x, y) of the point
mesh2 {
  vertex_vectors { 6, <-1, 0, 0>, <0, 0, 0>, <1, 0, 0>, <-1, 1.2, 0.3>, <0, 1.5, -0.2>, <1, 1.2, 0.3> }
  uv_vectors { 6, <0, 0>, <0.5, 0>, <1, 0>, <0, 1>, <0.5, 1>, <1, 1> }
  face_indices { 4, <0, 1, 4>, <0, 4, 3>, <1, 2, 5>, <1, 5, 4> }
  uv_indices { 4, <0, 1, 4>, <0, 4, 3>, <1, 2, 5>, <1, 5, 4> }
  uv_mapping
  texture { pigment { checker rgb <0.2, 0.6, 0.9>, rgb 1 scale 0.125 } finish { specular 0.3 } }
  rotate -15 * y scale 1.3 translate <0.2, 0, -1.6>
}
mesh2 {
  vertex_vectors { 4, <0, 0, 0>, <1.6, 0, 0>, <1.6, 1.4, 0>, <0, 1.4, 0> }
  uv_vectors { 4, <0, 0>, <2, 0>, <2, 2>, <0, 2> }
  face_indices { 2, <0, 1, 2>, <0, 2, 3> }
  uv_indices { 2, <0, 1, 2>, <0, 2, 3> }
  pigment { uv_mapping spiral2 5 color_map { [0 rgbf <1, 0.3, 0.2, 0.6>] [1 rgbf <0.2, 0.4, 1, 0.2>] } translate <1, 1, 0> }
  rotate 25 * y translate <2.4, 0, -1.8>
}
