// reflection { ... exponent E }: resultColour += reflec * Pow(rflCol, Reflect_Exp) (trace.cpp:1166-1168) - non-linear in the reflected
// ray's colour; mirrors facing each other (nested exponents), an exponent on one layer of a layered texture, with refraction behind it
#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 6 }
background { rgb <0.25, 0.35, 0.6> }
camera { location <0, 2.2, -7> look_at <0, 0.9, 0> angle 42 right x*16/9 }
light_source { <5, 8, -6> rgb <1, 0.95, 0.9> }
light_source { <-6, 5, -3> rgb <0.3, 0.35, 0.5> }
plane { y, 0 pigment { checker rgb <0.9,0.9,0.9>, rgb <0.15,0.2,0.3> } finish { ambient 0.1 diffuse 0.7 reflection { 0.35 exponent 0.6 } } }
sphere { <-1.6, 1, 0.5>, 1 pigment { rgb <0.9, 0.3, 0.2> } finish { ambient 0.05 diffuse 0.3 reflection { 0.1, 0.8 exponent 2.0 } } }
sphere { <1.5, 0.8, -0.3>, 0.8 pigment { rgb <0.2, 0.5, 0.9> } finish { ambient 0.05 diffuse 0.4 phong 0.6 reflection { 0.6 exponent 0.45 metallic } } }
box { <-0.5, 0, 1.6>, <0.7, 1.8, 1.9> rotate y*12
  texture { pigment { rgb <0.8, 0.8, 0.3> } finish { diffuse 0.3 reflection { 0.5 exponent 1.7 } } }
  texture { pigment { gradient y color_map { [0 rgbt <0.1,0.6,0.2,0.3>] [1 rgbt <0.1,0.6,0.2,0.95>] } scale 1.8 } finish { diffuse 0.2 reflection { 0.3 } } } }
sphere { <0.1, 0.55, -1.8>, 0.55 pigment { rgbf <0.9, 0.95, 1.0, 0.85> } finish { diffuse 0.05 specular 0.7 roughness 0.01 reflection { 0.05, 0.6 fresnel exponent 0.8 } } interior { ior 1.45 } }
