// normal_map: pattern-selected and blended normals, block-pattern normal lists, average normal_map, nested maps
#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 3 }
camera { location <0, 5, -11> look_at <0, 1.0, 0> angle 46 right x*16/9 }
light_source { <12, 18, -14> rgb <1, 1, 1> }
light_source { <-8, 5, -8> rgb <0.35, 0.35, 0.4> }
background { rgb <0.06, 0.08, 0.12> }
plane { y, 0 pigment { rgb <0.8, 0.8, 0.85> } normal { checker normal { bumps 0.6 scale 0.2 }, normal { ripples 0.5 frequency 2 scale 0.5 } scale 1.5 }
  finish { ambient 0.1 diffuse 0.7 specular 0.3 reflection 0.15 } }
sphere { <-4.0, 1.2, 0.5>, 1.2 pigment { rgb <0.9, 0.3, 0.25> }
  normal { gradient y normal_map { [0.2 dents 1.0 scale 0.3] [0.5 wrinkles 0.6 scale 0.4] [0.9 granite 0.7 scale 0.5] } scale 2.4 } finish { ambient 0.1 diffuse 0.7 phong 0.5 } }
sphere { <-1.3, 1.2, 0.5>, 1.2 pigment { rgb <0.3, 0.8, 0.4> }
  normal { average normal_map { [1 bumps 0.8 scale 0.2] [2 waves 0.7 frequency 3 scale 0.6] [1 quilted 0.6 scale 0.4] } rotate z*20 } finish { ambient 0.1 diffuse 0.7 specular 0.4 } }
sphere { <1.4, 1.2, 0.5>, 1.2 pigment { rgb <0.3, 0.4, 0.9> }
  normal { bozo normal_map { [0.4 bumps 0.7 scale 0.15] [0.6 marble normal_map { [0 dents 0.8 scale 0.2] [1 agate 0.6 scale 0.3] } scale 0.4] } scale 0.7 } finish { ambient 0.1 diffuse 0.7 phong 0.6 } }
sphere { <4.1, 1.2, 0.5>, 1.2 pigment { rgb <0.9, 0.8, 0.3> }
  normal { hexagon normal { bumps 0.8 scale 0.1 }, normal { dents 0.9 scale 0.2 }, normal { wrinkles 0.7 scale 0.3 } scale 0.4 rotate x*90 } finish { ambient 0.1 diffuse 0.7 specular 0.3 } }
