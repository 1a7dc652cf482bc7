// text objects (TrueType glyphs, truetype.cpp:2392-2976): plain, sheared / rotated, as operands of difference and intersection,
// clipped, and transparent with an interior (wall and face normals under refraction and in filtered shadows)
#version 3.7;
global_settings { assumed_gamma 1.0 max_trace_level 6 }
camera { location <0.5, 2.2, -8.5> look_at <0.3, 0.8, 0> angle 46 }
light_source { <-6, 9, -10> rgb <1, 0.95, 0.9> }
light_source { <7, 4, -6> rgb <0.3, 0.35, 0.5> }
plane { y, -0.6 pigment { checker rgb <0.85, 0.85, 0.8>, rgb <0.3, 0.35, 0.45> scale 1.5 } finish { reflection 0.15 } }
text { internal 1 "B200 &Qg" 0.35, 0 pigment { rgb <0.9, 0.5, 0.15> } finish { phong 0.6 } translate <-3.6, 1.6, 0.5> }
text { internal 2 "Ray%" 0.6, <0.05, 0, 0> pigment { bozo color_map { [0 rgb <0.2, 0.6, 0.3>] [1 rgb <0.9, 0.9, 0.2>] } scale 0.2 }
       matrix <1, 0, 0,  0.35, 1, 0,  0, 0, 1,  0, 0, 0> rotate <-15, 25, 0> scale 1.3 translate <0.8, 1.4, 1> }
difference {
  box { <-3.6, -0.5, -0.3>, <-0.4, 0.8, 0.3> }
  text { internal 1 "CSG" 1, 0 scale <1.35, 1.35, 1> translate <-3.45, -0.3, -0.5> }
  pigment { rgb <0.55, 0.6, 0.9> } finish { specular 0.4 }
  rotate 12 * y
}
intersection {
  text { internal 3 "8@" 0.8, 0 scale 1.6 translate <0, -0.4, -0.4> }
  sphere { <0.9, 0.25, 0>, 0.95 }
  pigment { rgb <0.9, 0.25, 0.3> } finish { phong 0.4 }
  translate <0.3, 0, -0.5>
}
text { internal 1 "glass" 0.5, 0 pigment { rgbf <0.85, 1, 0.9, 0.8> } finish { specular 0.5 reflection 0.1 } interior { ior 1.45 }
       scale 1.1 rotate 20 * x translate <1.7, -0.45, -2.2> }
text { internal 2 "clip" 0.4, 0 pigment { rgb <0.8, 0.8, 0.2> } clipped_by { plane { <1, 0.6, 0>, 1.15 } } translate <-2.6, -0.5, -2.6> }
