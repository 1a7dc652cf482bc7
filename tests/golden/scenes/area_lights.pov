// area lights (SURVEY 8f rank 2): rectangular adaptive / non-adaptive, circular + orient, a spot area light, opaque and
// filtered (glass) shadow casters; no jitter (jitter draws from the per-thread RNG and is excluded from parity)
#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 4 }
camera { location <0, 6, -11> look_at <0, 0.8, 0> angle 48 right x*16/9 }
light_source { <8, 10, -6> rgb <0.8, 0.75, 0.7> area_light <3, 0, 0>, <0, 0, 3>, 5, 5 adaptive 1 }
light_source { <-7, 8, -3> rgb <0.35, 0.4, 0.5> area_light <2, 0, 0>, <0, 2, 0>, 4, 3 }
light_source { <0, 9, 6> rgb <0.4, 0.35, 0.3> area_light <2.5, 0, 0>, <0, 0, 2.5>, 7, 7 adaptive 0 circular orient }
light_source { <3, 7, -9> rgb <0.5, 0.5, 0.4> spotlight point_at <1, 0, 0> radius 12 falloff 20 tightness 2 area_light <1.5, 0, 0>, <0, 1.5, 0>, 3, 3 adaptive 2 }
background { rgb <0.05, 0.05, 0.08> }
plane { y, 0 pigment { checker rgb <0.85, 0.85, 0.85>, rgb <0.55, 0.6, 0.65> } finish { ambient 0.08 diffuse 0.8 } }
sphere { <-3.0, 1.0, 0.0>, 1.0 pigment { rgb <0.9, 0.3, 0.25> } finish { ambient 0.1 diffuse 0.7 phong 0.5 } }
box { <-0.8, 0, -0.8>, <0.8, 1.8, 0.8> pigment { rgb <0.3, 0.7, 0.4> } finish { ambient 0.1 diffuse 0.7 } rotate y*30 translate <0.3, 0, 1.0> }
cylinder { <3.2, 0, 0.5>, <3.2, 2.2, 0.5>, 0.6 pigment { rgb <0.3, 0.4, 0.9> } finish { ambient 0.1 diffuse 0.7 specular 0.4 } }
sphere { <1.2, 0.7, -2.6>, 0.7 pigment { rgbf <0.9, 1.0, 0.9, 0.8> } finish { ambient 0.02 diffuse 0.2 specular 0.5 roughness 0.02 } interior { ior 1.4 } }
torus { 0.9, 0.25 pigment { rgb <0.9, 0.8, 0.3> } finish { ambient 0.1 diffuse 0.7 } rotate x*20 translate <-1.5, 0.6, -3.0> }
