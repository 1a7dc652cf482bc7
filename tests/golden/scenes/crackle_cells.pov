// crackle (metrics 2 / 1 / 3, form, offset, solid, repeat) and cells pigments, crackle as a normal (SURVEY 8f rank 2)
#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 3 }
camera { location <0, 5, -11> look_at <0, 1.0, 0> angle 46 right x*16/9 }
light_source { <12, 18, -14> rgb <1, 1, 1> }
background { rgb <0.06, 0.08, 0.12> }
plane { y, 0 pigment { crackle color_map { [0 rgb <0.1, 0.1, 0.1>] [0.08 rgb <0.7, 0.65, 0.6>] [1 rgb <0.9, 0.85, 0.8>] } scale 0.8 } finish { ambient 0.1 diffuse 0.7 } }
sphere { <-4.4, 1.1, 0.5>, 1.1 pigment { crackle metric 1 form <-1, 1, 0> color_map { [0 rgb <0.9, 0.2, 0.2>] [1 rgb <1, 1, 0.6>] } scale 0.5 } finish { ambient 0.1 diffuse 0.7 } }
sphere { <-2.2, 1.1, 0.5>, 1.1 pigment { crackle metric 3 offset 0.2 color_map { [0 rgb <0.1, 0.3, 0.8>] [1 rgb <0.9, 0.95, 1>] } scale 0.45 rotate x*30 } finish { ambient 0.1 diffuse 0.7 phong 0.4 } }
sphere { <0.0, 1.1, 0.5>, 1.1 pigment { crackle solid color_map { [0 rgb <0.2, 0.6, 0.2>] [0.5 rgb <0.9, 0.8, 0.2>] [1 rgb <0.8, 0.2, 0.5>] } scale 0.4 } finish { ambient 0.1 diffuse 0.7 } }
sphere { <2.2, 1.1, 0.5>, 1.1 pigment { crackle repeat <2, 3, 2> form <1, 0, 0> color_map { [0 rgb <1, 1, 1>] [1 rgb <0.2, 0.1, 0.4>] } scale 0.35 } finish { ambient 0.1 diffuse 0.7 } }
sphere { <4.4, 1.1, 0.5>, 1.1 pigment { cells color_map { [0 rgb <0.1, 0.1, 0.5>] [0.5 rgb <0.2, 0.8, 0.6>] [1 rgb <1, 0.9, 0.4>] } scale 0.3 } finish { ambient 0.1 diffuse 0.7 } }
box { <-1.5, 0, -3.8>, <1.5, 1.0, -2.6> pigment { rgb <0.8, 0.75, 0.7> } normal { crackle 0.8 scale 0.3 } finish { ambient 0.1 diffuse 0.7 specular 0.3 } }
