// triangles, smooth triangles and polygons (SURVEY 8f rank 1), alone and inside CSG
#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 5 }
camera { location <0, 4.5, -11> look_at <0, 1.2, 0> angle 48 right x*16/9 }
light_source { <12, 18, -14> rgb <1, 1, 1> }
light_source { <-9, 7, -6> rgb <0.35, 0.35, 0.45> }
background { rgb <0.05, 0.07, 0.12> }
plane { y, 0 pigment { checker rgb <0.9, 0.9, 0.9>, rgb <0.2, 0.25, 0.3> } finish { ambient 0.1 diffuse 0.7 reflection 0.15 } }
triangle { <-4.5, 0.2, 1.0>, <-2.6, 0.3, 1.6>, <-3.7, 2.6, 0.8> pigment { rgb <0.9, 0.3, 0.2> } finish { ambient 0.1 diffuse 0.7 phong 0.4 } }
triangle { <-1.0, 0.1, -2.0>, <0.2, 0.15, -2.6>, <-0.3, 1.4, -1.9> pigment { rgb <0.2, 0.8, 0.3> } finish { ambient 0.1 diffuse 0.6 reflection 0.3 }
  rotate y*25 translate <0.4, 0, 0.3> }
smooth_triangle { <1.2, 0.2, 0.5>, <-0.4, 0.1, -1>, <3.4, 0.3, 1.3>, <0.5, 0.2, -1>, <2.1, 2.8, 0.6>, <0.0, 0.8, -0.7>
  pigment { rgb <0.3, 0.4, 0.95> } finish { ambient 0.1 diffuse 0.6 specular 0.6 roughness 0.02 } }
smooth_triangle { <-2.0, 0.3, 3.5>, <-0.3, 0.3, -1>, <0.2, 0.4, 4.1>, <0.4, 0.2, -1>, <-1.0, 2.9, 3.9>, <0.0, 1.0, -0.5>
  pigment { rgbt <0.9, 0.8, 0.3, 0.4> } finish { ambient 0.1 diffuse 0.6 phong 0.8 phong_size 40 } scale <1.1, 0.9, 1> }
polygon { 5, <0, 0>, <1.6, 0>, <2.0, 1.2>, <0.8, 2.0>, <-0.3, 1.1> pigment { rgb <0.9, 0.7, 0.2> } finish { ambient 0.1 diffuse 0.7 }
  rotate x*-15 rotate y*-30 translate <3.2, 0.2, -1.5> }
// polygon with a hole (two closed sub-polygons)
polygon { 10, <0, 0>, <2.4, 0>, <2.4, 2.2>, <0, 2.2>, <0, 0>, <0.6, 0.5>, <1.8, 0.5>, <1.8, 1.6>, <0.6, 1.6>, <0.6, 0.5>
  pigment { rgb <0.7, 0.3, 0.8> } finish { ambient 0.1 diffuse 0.7 specular 0.2 } rotate y*20 translate <-6.4, 0.1, 3.0> }
// a polygon in 3D (projected by Compute_Polygon)
polygon { 4, <4.2, 0.2, 2.5>, <6.0, 0.3, 3.5>, <5.8, 2.3, 3.9>, <4.0, 2.2, 2.9> pigment { gradient y color_map { [0 rgb <1, 0.2, 0.2>] [1 rgb <0.2, 0.2, 1>] } scale 2.5 }
  finish { ambient 0.15 diffuse 0.7 } }
// patches clipped by a solid, and a union holding a triangle
triangle { <-4.4, 0.1, -2.5>, <-1.9, 0.1, -3.0>, <-3.1, 2.8, -2.2> clipped_by { sphere { <-3.1, 1.0, -2.6>, 1.1 } }
  pigment { rgb <0.2, 0.9, 0.9> } finish { ambient 0.1 diffuse 0.7 } }
union { triangle { <2.0, 0.1, -3.5>, <3.6, 0.1, -3.0>, <2.8, 1.9, -3.2> } sphere { <2.8, 2.2, -3.2>, 0.35 }
  pigment { rgb <0.95, 0.5, 0.1> } finish { ambient 0.1 diffuse 0.7 phong 0.5 } }
