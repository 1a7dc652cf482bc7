// normal perturbation (SURVEY 8f rank 1): bumps, dents, wrinkles, ripples, waves, quilted, pattern normals with and
// without slope maps, warped normals, bump_size scaling, perturbed reflection / refraction, layered textures
#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 5 }
camera { location <0, 5.5, -12> look_at <0, 1.0, 0> angle 46 right x*16/9 }
light_source { <14, 20, -16> rgb <1, 1, 1> }
light_source { <-10, 8, -5> rgb <0.3, 0.3, 0.4> }
background { rgb <0.1, 0.12, 0.2> }
plane { y, 0 pigment { rgb <0.55, 0.6, 0.7> } normal { ripples 0.6 frequency 1.5 phase 0.2 scale 1.7 }
  finish { ambient 0.1 diffuse 0.6 reflection 0.35 specular 0.4 roughness 0.02 } }
sphere { <-5.0, 1.0, 0.5>, 1.0 pigment { rgb <0.9, 0.3, 0.25> } normal { bumps 0.8 scale 0.25 } finish { ambient 0.1 diffuse 0.7 phong 0.6 } }
sphere { <-2.5, 1.0, 0.5>, 1.0 pigment { rgb <0.3, 0.8, 0.3> } normal { dents 1.2 scale 0.4 } finish { ambient 0.1 diffuse 0.7 specular 0.5 roughness 0.03 } }
sphere { <0.0, 1.0, 0.5>, 1.0 pigment { rgb <0.3, 0.4, 0.9> } normal { wrinkles 0.7 scale 0.6 rotate y*30 } finish { ambient 0.1 diffuse 0.6 reflection 0.25 } }
sphere { <2.5, 1.0, 0.5>, 1.0 pigment { rgb <0.9, 0.8, 0.3> } normal { waves 0.9 frequency 3 scale 0.8 translate <0.3, 0.2, 0> } finish { ambient 0.1 diffuse 0.7 phong 0.8 phong_size 60 } }
sphere { <5.0, 1.0, 0.5>, 1.0 pigment { rgb <0.8, 0.4, 0.8> } normal { quilted 0.8 control0 0.2 control1 0.7 scale 0.45 } finish { ambient 0.1 diffuse 0.7 specular 0.3 } }
box { <-6.0, 0, 3.0>, <-4.0, 2.0, 4.5> pigment { rgb <0.8, 0.8, 0.8> } normal { granite 0.6 scale 0.8 } finish { ambient 0.1 diffuse 0.7 } }
box { <-3.5, 0, 3.0>, <-1.5, 2.0, 4.5> pigment { rgb <0.9, 0.6, 0.4> } normal { gradient x 0.8 slope_map { [0 <0, 1>] [0.5 <1, 1>] [0.5 <1, -1>] [1 <0, -1>] } scale 0.4 }
  finish { ambient 0.1 diffuse 0.7 specular 0.4 } }
box { <-1.0, 0, 3.0>, <1.0, 2.0, 4.5> pigment { rgb <0.5, 0.8, 0.8> } normal { agate 0.7 scale 0.5 } finish { ambient 0.1 diffuse 0.7 } }
box { <1.5, 0, 3.0>, <3.5, 2.0, 4.5> pigment { rgb <0.7, 0.7, 0.4> } normal { marble 0.9 turbulence 0.6 sine_wave scale 0.5 } finish { ambient 0.1 diffuse 0.7 phong 0.3 } }
box { <4.0, 0, 3.0>, <6.0, 2.0, 4.5> pigment { rgb <0.6, 0.5, 0.9> } normal { bozo 1.5 bump_size 1.5 no_bump_scale scale <0.3, 0.6, 0.3> }
  finish { ambient 0.1 diffuse 0.7 specular 0.5 roughness 0.05 } }
// glass with a perturbed surface: refraction and total internal reflection use the perturbed top normal
sphere { <-1.3, 0.9, -3.5>, 0.9 pigment { rgbf <0.9, 1.0, 0.95, 0.85> } normal { bumps 0.3 scale 0.2 }
  finish { ambient 0.02 diffuse 0.2 specular 0.6 roughness 0.01 reflection 0.1 } interior { ior 1.45 caustics 0.8 } }
// layered texture: two layers with different normals
cylinder { <1.8, 0, -3.5>, <1.8, 1.8, -3.5>, 0.8
  texture { pigment { rgb <0.8, 0.2, 0.2> } normal { dents 0.8 scale 0.3 } finish { ambient 0.1 diffuse 0.6 reflection 0.2 } }
  texture { pigment { bozo color_map { [0.4 rgbt <1, 1, 1, 1>] [0.6 rgbt <0.2, 0.2, 0.9, 0.2>] } scale 0.4 } normal { onion 0.7 scale 0.3 } finish { ambient 0.1 diffuse 0.6 phong 0.5 reflection 0.15 } } }
