// image_map pigments (imageutil.cpp): planar / spherical / cylindrical / torus / angular mapping, interpolate 2 / 3 / 4, once,
// an RGBA image (transparent texels let the layer below and the shadow through), transmit all, inside a pigment_map
#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 5 }
background { rgb <0.2, 0.25, 0.4> }
camera { location <0.02, 3.0, -8.5> look_at <0, 1.0, 0> angle 42 right x*16/9 }
light_source { <5, 9, -6> rgb 1 }
light_source { <-6, 5, -4> rgb <0.3, 0.3, 0.35> }
plane { y, -0.0078125 pigment { image_map { ppm "img_ramp.ppm" interpolate 2 } rotate x*90 scale 3 } finish { ambient 0.15 diffuse 0.7 } }
sphere { 0, 1 pigment { image_map { ppm "img_ramp.ppm" map_type 1 interpolate 4 } } finish { ambient 0.1 diffuse 0.7 phong 0.4 } rotate y*30 translate <-3.2, 1, 0.5> }
cylinder { <0, 0, 0>, <0, 1, 0>, 0.7 pigment { image_map { ppm "img_ramp.ppm" map_type 2 interpolate 3 once } } finish { ambient 0.1 diffuse 0.7 } scale <1, 2, 1> translate <-1.3, 0, 0.8> }
torus { 0.8, 0.3 pigment { image_map { ppm "img_ramp.ppm" map_type 5 } } finish { ambient 0.1 diffuse 0.7 } rotate x*-50 translate <0.6, 1.1, 0> }
sphere { 0, 0.9 pigment { image_map { png "img_disc.png" map_type 7 interpolate 2 transmit all 0.3 } } finish { ambient 0.1 diffuse 0.7 } translate <2.4, 0.9, 0.6> }
box { <0, 0, 0>, <1, 1, 0.05>
  texture { pigment { rgb <0.9, 0.8, 0.2> } finish { ambient 0.1 diffuse 0.6 } }
  texture { pigment { image_map { png "img_disc.png" once interpolate 2 } } finish { ambient 0.1 diffuse 0.6 } }
  scale <2, 2, 1> rotate y*-25 translate <1.6, 0.2, -2.2> }
box { <-0.5, 0, -0.5>, <0.5, 1, 0.5>
  pigment { gradient y pigment_map { [0.2 image_map { ppm "img_ramp.ppm" } scale 0.5] [0.8 rgb <0.2, 0.6, 0.3>] } }
  finish { ambient 0.1 diffuse 0.7 } rotate y*40 translate <-0.6, 0, -2.6> }
