// poly / cubic / quartic objects of order <= 4 (polynomial.cpp): closed-form and sturm, in CSG, clipped, transformed
#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 4 }
camera { location <0, 5, -11> look_at <0, 1.2, 0> angle 46 right x*16/9 }
light_source { <12, 18, -14> rgb <1, 1, 1> }
light_source { <-9, 6, -6> rgb <0.3, 0.3, 0.4> }
background { rgb <0.06, 0.08, 0.12> }
plane { y, 0 pigment { checker rgb <0.9, 0.9, 0.9>, rgb <0.3, 0.35, 0.4> } finish { ambient 0.1 diffuse 0.7 } }
// torus as a quartic (shapesq.inc Torus_40_12 style)
quartic { <1, 0, 0, 0, 2, 0, 0, 2, 0, -2.1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 2, 0, 1.9, 0, 0, 0, 0, 1, 0, -2.1, 0, 0.9025>
  sturm pigment { rgb <0.9, 0.4, 0.3> } finish { ambient 0.1 diffuse 0.7 phong 0.5 } rotate x*35 translate <-4.2, 1.4, 0.5> }
// lemniscate of Gerono revolved (quartic, closed form solver)
quartic { <1, 0, 0, 0, 0, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0>
  bounded_by { sphere { <0, 0, 0>, 2 } } pigment { rgb <0.3, 0.8, 0.4> } finish { ambient 0.1 diffuse 0.7 specular 0.4 } scale 1.3 rotate y*40 translate <-1.2, 1.2, 1.5> }
// a cubic saddle clipped to a box
cubic { <0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0, -1, 0>
  clipped_by { box { <-1.2, -1.2, -1.2>, <1.2, 1.2, 1.2> } } bounded_by { clipped_by }
  pigment { rgb <0.3, 0.5, 0.9> } finish { ambient 0.1 diffuse 0.7 } scale 0.9 rotate <20, 30, 0> translate <1.8, 1.4, 0.8> }
// quadric surface given as a poly of order 2 and a plane as order 1
poly { 2, <1, 0, 0, 0, 2, 0, 0, 0.5, 0, -1> pigment { rgb <0.9, 0.8, 0.3> } finish { ambient 0.1 diffuse 0.7 phong 0.4 } translate <4.6, 1.0, 0.5> }
intersection { poly { 4, <0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, -0.8> }
  plane { <0.3, 1, 0.2>, 0.2 } pigment { rgb <0.8, 0.3, 0.8> } finish { ambient 0.1 diffuse 0.7 } translate <0.5, 0.9, -3.0> }
// glass piriform (quartic with refraction)
quartic { <4, 0, 0, -4, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0>
  sturm pigment { rgbf <0.9, 0.95, 1, 0.8> } finish { ambient 0.02 diffuse 0.2 specular 0.5 roughness 0.02 } interior { ior 1.3 }
  scale <1.6, 1.6, 1.6> rotate z*90 translate <-3.0, 0.1, -3.0> }
