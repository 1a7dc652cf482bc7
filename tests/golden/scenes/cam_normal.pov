// camera { normal { ... } }: the primary ray direction is perturbed like a surface normal at (x0, y0, 0) (tracepixel.cpp:917-924)
#version 3.7;
global_settings { assumed_gamma 1.0 }
camera { location <0, 2.5, -8> look_at <0, 0.8, 0> angle 45 normal { bumps 0.25 scale 0.12 } }
light_source { <-6, 9, -7> rgb 1 }
plane { y, 0 pigment { checker rgb <0.85, 0.85, 0.8>, rgb <0.3, 0.35, 0.45> } }
sphere { <-1.8, 1, 0.5>, 1 pigment { rgb <0.9, 0.5, 0.3> } finish { phong 0.7 reflection 0.2 } }
box { <0.4, 0, -0.4>, <2.2, 1.8, 1.4> pigment { rgb <0.4, 0.6, 0.9> } rotate 20 * y }
