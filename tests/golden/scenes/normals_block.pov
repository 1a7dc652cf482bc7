// block patterns as normals WITHOUT a normal_map: sampled on the pyramid like any other pattern (normal.cpp:880-905)
#version 3.7;
global_settings { assumed_gamma 1.0 }
camera { location <0, 3.5, -8> look_at <0, 0.8, 0> angle 45 }
light_source { <-6, 9, -7> rgb 1 }
light_source { <7, 5, -5> rgb <0.3, 0.35, 0.5> }
plane { y, 0 pigment { rgb <0.8, 0.8, 0.75> } normal { hexagon 0.6 scale 0.7 } finish { specular 0.3 } }
sphere { <-2.6, 1, 0.5>, 1 pigment { rgb <0.9, 0.5, 0.3> } normal { checker 0.8 scale 0.25 } finish { phong 0.7 } }
box { <-0.9, 0, -0.4>, <0.9, 1.8, 1.4> pigment { rgb <0.4, 0.6, 0.9> } normal { brick 0.5 scale 0.12 } finish { specular 0.4 } rotate 20 * y }
cylinder { <2.7, 0, 0.3>, <2.7, 2, 0.3>, 0.8 pigment { rgb <0.5, 0.8, 0.4> } normal { hexagon 1.2 scale 0.3 rotate 40 * x } finish { phong 0.5 reflection 0.15 } }
