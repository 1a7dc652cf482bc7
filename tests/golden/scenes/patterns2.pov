// the remaining cheap pigment patterns (SURVEY 8f rank 2): brick, hexagon, wood, leopard, spherical, boxed, radial,
// cylindrical, planar, dents, ripples, waves, quilted, bumps - with waveforms, turbulence and transforms
#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 4 number_of_waves 7 }
camera { location <0, 6.5, -13> look_at <0, 1.0, 0.5> angle 46 right x*16/9 }
light_source { <10, 20, -15> rgb <1, 1, 1> }
light_source { <-12, 9, -4> rgb <0.3, 0.3, 0.35> }
background { rgb <0.08, 0.09, 0.14> }
plane { y, 0 pigment { hexagon rgb <0.9, 0.85, 0.8>, rgb <0.3, 0.35, 0.5>, rgb <0.6, 0.3, 0.3> scale 0.9 rotate y*15 } finish { ambient 0.1 diffuse 0.7 } }
box { <-7.0, 0, 4.0>, <7.0, 3.0, 4.6> pigment { brick rgb <0.8, 0.8, 0.75>, rgb <0.7, 0.25, 0.2> brick_size <0.9, 0.35, 0.5> mortar 0.06 } finish { ambient 0.1 diffuse 0.7 } }
sphere { <-5.5, 1.0, 1.0>, 1.0 pigment { wood turbulence 0.08 color_map { [0 rgb <0.6, 0.4, 0.2>] [0.6 rgb <0.4, 0.25, 0.1>] [1 rgb <0.3, 0.15, 0.05>] } scale 0.25 rotate x*20 } finish { ambient 0.1 diffuse 0.7 phong 0.3 } }
sphere { <-3.3, 1.0, 1.0>, 1.0 pigment { leopard color_map { [0 rgb <1, 0.9, 0.3>] [0.5 rgb <0.8, 0.4, 0.1>] [1 rgb <0.1, 0.05, 0>] } scale 0.12 } finish { ambient 0.1 diffuse 0.7 } }
sphere { <-1.1, 1.0, 1.0>, 1.0 pigment { spherical color_map { [0 rgb <0.1, 0.1, 0.6>] [1 rgb <1, 1, 0.6>] } scale 1.4 translate <-1.1, 1.6, 0.6> } finish { ambient 0.1 diffuse 0.7 } }
sphere { <1.1, 1.0, 1.0>, 1.0 pigment { boxed triangle_wave frequency 3 color_map { [0 rgb <0.2, 0.7, 0.3>] [1 rgb <0.9, 0.9, 0.2>] } scale 1.2 translate <1.1, 1.0, 1.0> } finish { ambient 0.1 diffuse 0.7 } }
sphere { <3.3, 1.0, 1.0>, 1.0 pigment { radial frequency 6 sine_wave color_map { [0 rgb <0.9, 0.2, 0.2>] [1 rgb <0.95, 0.95, 0.95>] } translate <3.3, 0, 1.0> } finish { ambient 0.1 diffuse 0.7 specular 0.3 } }
sphere { <5.5, 1.0, 1.0>, 1.0 pigment { cylindrical scallop_wave color_map { [0 rgb <0.1, 0.5, 0.6>] [1 rgb <0.9, 0.95, 1>] } scale 0.8 rotate z*35 translate <5.5, 1.0, 1.0> } finish { ambient 0.1 diffuse 0.7 } }
cylinder { <-5.0, 0, -2.5>, <-5.0, 1.6, -2.5>, 0.8 pigment { planar cubic_wave color_map { [0 rgb <0.2, 0.2, 0.2>] [1 rgb <1, 0.6, 0.1>] } scale 1.7 } finish { ambient 0.1 diffuse 0.7 } }
cylinder { <-2.5, 0, -2.5>, <-2.5, 1.6, -2.5>, 0.8 pigment { dents poly_wave 0.6 color_map { [0 rgb <0.9, 0.9, 0.9>] [1 rgb <0.1, 0.2, 0.5>] } scale 0.3 } finish { ambient 0.1 diffuse 0.7 } }
cylinder { <0.0, 0, -2.5>, <0.0, 1.6, -2.5>, 0.8 pigment { ripples frequency 2 phase 0.3 color_map { [0 rgb <0.1, 0.3, 0.8>] [1 rgb <0.8, 0.95, 1>] } scale 0.5 } finish { ambient 0.1 diffuse 0.7 } }
cylinder { <2.5, 0, -2.5>, <2.5, 1.6, -2.5>, 0.8 pigment { waves frequency 1.5 color_map { [0 rgb <0.6, 0.1, 0.5>] [1 rgb <1, 0.9, 0.6>] } scale 0.4 } finish { ambient 0.1 diffuse 0.7 } }
cylinder { <5.0, 0, -2.5>, <5.0, 1.6, -2.5>, 0.8 pigment { quilted control0 0.3 control1 0.8 color_map { [0 rgb <0.2, 0.6, 0.2>] [1 rgb <1, 1, 0.8>] } scale 0.5 } finish { ambient 0.1 diffuse 0.7 } }
box { <-1.0, 0, -5.5>, <1.0, 0.8, -4.5> pigment { bumps color_map { [0 rgb <0.3, 0.2, 0.1>] [1 rgb <0.9, 0.8, 0.6>] } scale 0.2 } normal { ripples 0.5 scale 0.3 } finish { ambient 0.1 diffuse 0.7 } }
