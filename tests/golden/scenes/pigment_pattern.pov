// pigment_pattern: the greyscale of a pigment as the pattern value of another pigment and of a normal (pattern.cpp:7974-7990)
#version 3.7;
global_settings { assumed_gamma 1.0 }
camera { location <0, 3.5, -8> look_at <0, 0.8, 0> angle 45 }
light_source { <-6, 9, -7> rgb 1 }
light_source { <7, 5, -5> rgb <0.3, 0.35, 0.5> }
#declare Bricks = pigment { brick rgb 0.1, rgb 0.9 brick_size <0.5, 0.25, 0.3> mortar 0.04 }
#declare Blotch = pigment { bozo color_map { [0 rgb 0] [0.45 rgb 0.2] [0.55 rgb 0.8] [1 rgb 1] } scale 0.4 }
plane { y, 0 pigment { pigment_pattern { checker rgb 0, rgb 1 scale 0.8 } color_map { [0 rgb <0.7, 0.3, 0.2>] [1 rgb <0.9, 0.9, 0.7>] } }
        normal { pigment_pattern { Bricks rotate 90 * x } 0.8 } }
sphere { <-2.5, 1, 0.5>, 1 pigment { pigment_pattern { Blotch } color_map { [0 rgb <0.1, 0.2, 0.6>] [0.5 rgb <0.9, 0.8, 0.3>] [1 rgb <0.8, 0.2, 0.2>] } sine_wave frequency 2 } finish { phong 0.6 } }
box { <-0.9, 0, -0.4>, <0.9, 1.8, 1.4> pigment { rgb <0.75, 0.5, 0.4> } normal { pigment_pattern { Bricks } 1.0 slope_map { [0 <0, 0>] [0.5 <0.5, 1>] [1 <1, 0>] } } finish { specular 0.3 } rotate 20 * y }
cylinder { <2.7, 0, 0.3>, <2.7, 2, 0.3>, 0.8
  pigment { pigment_pattern { pigment_pattern { Blotch } color_map { [0 rgb 0] [0.5 rgb 1] [1 rgb 0] } } pigment_map { [0 Bricks scale 0.5] [1 rgb <0.3, 0.7, 0.4>] } } }
