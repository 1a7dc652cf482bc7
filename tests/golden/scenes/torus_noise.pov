// tori (closed form + sturm) with granite / bozo / marble / agate / wrinkles / gradient / onion pigments (config-4 flavour)
#version 3.7;
global_settings { assumed_gamma 1 }
background { rgb <0.1, 0.1, 0.14> }
camera { location <0, 4.5, -8> look_at <0, 0.6, 0> angle 45 right x*16/9 }
light_source { <5, 10, -8> rgb 1 }
light_source { <-7, 6, 2> rgb <0.4, 0.4, 0.5> spotlight point_at <0, 0, 0> radius 20 falloff 35 tightness 2 }
plane { y, -0.0078125 pigment { granite color_map { [0 rgb <0.2,0.2,0.25>] [0.5 rgb <0.6,0.55,0.5>] [1 rgb 1] } scale 2 } finish { ambient 0.1 diffuse 0.7 } }
torus { 1.0, 0.3 pigment { bozo color_map { [0 rgb <1,0.2,0.1>] [0.4 rgb <1,0.8,0.1>] [1 rgb <0.1,0.2,1>] } scale 0.3 } finish { phong 0.5 } rotate x*25 translate <-2.6, 0.9, 0.5> }
torus { 0.9, 0.25 sturm pigment { marble turbulence 0.6 color_map { [0 rgb 1] [0.7 rgb <0.3,0.3,0.3>] [1 rgb 0] } scale 0.5 } finish { specular 0.4 } rotate <70, 20, 0> translate <0, 1.1, 0> }
torus { 0.8, 0.35 pigment { agate color_map { [0 rgb <0.4,0.2,0.1>] [1 rgb <1,0.9,0.7>] } scale 0.6 } rotate z*40 translate <2.6, 1.0, 0.2> }
torus { 0.6, 0.2 pigment { wrinkles color_map { [0 rgb <0.1,0.4,0.2>] [1 rgb <0.9,1,0.8>] } scale 0.4 frequency 2 sine_wave } rotate x*90 translate <-1.2, 0.7, -2.2> }
torus { 0.6, 0.2 sturm pigment { gradient y color_map { [0 rgb <1,0,0>] [0.5 rgb <0,1,0>] [1 rgb <0,0,1>] } scale 0.5 triangle_wave } translate <1.3, 0.25, -2.3> }
sphere { <0, 0.6, -2.6>, 0.6 pigment { onion color_map { [0 rgb <0.9,0.9,0.2>] [1 rgb <0.5,0.1,0.5>] } scale 0.25 warp { turbulence 0.3 octaves 3 } } finish { phong 0.3 } }
sphere { <3.2, 0.5, -1.8>, 0.5 pigment { spotted color_map { [0 rgb 0.1] [1 rgb <0.9,0.6,0.3>] } scale 0.2 } }
sphere { <-3.3, 0.5, -1.6>, 0.5 pigment { bozo noise_generator 3 color_map { [0 rgb <0,0.3,0.6>] [1 rgb 1] } scale 0.15 } }
sphere { <-3.4, 0.5, 2.6>, 0.5 pigment { granite noise_generator 1 color_map { [0 rgb <0.3,0,0>] [1 rgb <1,1,0.6>] } scale 0.8 } }
