// iridescence (thin-film interference, trace.cpp:2486-2518): on the light terms and on reflections, with turbulence, layered
#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 4 irid_wavelength rgb <0.70, 0.52, 0.48> }
camera { location <0, 4.5, -10> look_at <0, 1.0, 0> angle 46 right x*16/9 }
light_source { <12, 18, -14> rgb <1, 1, 1> }
light_source { <-8, 6, -6> rgb <0.4, 0.4, 0.45> }
background { rgb <0.06, 0.08, 0.12> }
plane { y, 0 pigment { checker rgb <0.9, 0.9, 0.9>, rgb <0.3, 0.3, 0.35> } finish { ambient 0.1 diffuse 0.7 } }
sphere { <-3.5, 1.2, 0.5>, 1.2 pigment { rgb <0.2, 0.2, 0.25> } finish { ambient 0.05 diffuse 0.5 specular 0.8 roughness 0.02 irid { 0.35 thickness 0.5 turbulence 0.4 } } }
sphere { <-0.6, 1.2, 0.5>, 1.2 pigment { rgb <0.6, 0.6, 0.65> } finish { ambient 0.05 diffuse 0.3 phong 0.6 reflection 0.5 irid { 0.5 thickness 0.3 } } }
sphere { <2.3, 1.2, 0.5>, 1.2 pigment { rgbf <0.95, 1, 0.95, 0.85> } finish { ambient 0.02 diffuse 0.1 specular 0.6 roughness 0.01 reflection 0.15 irid { 0.4 thickness 0.8 turbulence 0.2 } } interior { ior 1.4 } }
box { <4.0, 0, -0.5>, <5.6, 1.8, 1.2>
  texture { pigment { rgb <0.8, 0.3, 0.2> } finish { ambient 0.1 diffuse 0.6 irid { 0.3 thickness 0.4 } } }
  texture { pigment { bozo color_map { [0.4 rgbt <1, 1, 1, 1>] [0.7 rgbt <0.2, 0.3, 0.9, 0.3>] } scale 0.3 } finish { ambient 0.1 diffuse 0.5 reflection 0.2 irid { 0.6 thickness 0.25 turbulence 0.6 } } }
  rotate y*25 }
