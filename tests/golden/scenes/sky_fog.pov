// sky_sphere (layered pigments, transformed) and fog (constant + ground fog with turbulence), seen directly, in
// reflections and through glass (SURVEY 8f rank 2)
#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 5 }
camera { location <0, 3.0, -12> look_at <0, 1.5, 0> angle 50 right x*16/9 }
light_source { <15, 25, -20> rgb <1, 0.95, 0.9> }
background { rgb <0.02, 0.02, 0.05> }
sky_sphere {
  pigment { gradient y color_map { [0 rgb <0.9, 0.7, 0.5>] [0.3 rgb <0.4, 0.6, 0.9>] [1 rgb <0.05, 0.15, 0.5>] } scale 2 translate -1 }
  pigment { bozo turbulence 0.6 color_map { [0.45 rgbt <1, 1, 1, 1>] [0.7 rgbt <1, 1, 1, 0.3>] [1 rgbt <0.9, 0.9, 0.9, 0.1>] } scale <0.4, 0.12, 0.4> }
  emission rgb <1.0, 0.95, 0.9>
  rotate z*5
}
fog { distance 45 colour rgbt <0.7, 0.75, 0.8, 0.15> }
fog { fog_type 2 distance 12 colour rgbf <0.85, 0.85, 0.9, 0.1> fog_offset 0.3 fog_alt 0.8 turbulence 0.4 turb_depth 0.3 up <0, 1, 0.1> }
plane { y, 0 pigment { checker rgb <0.8, 0.8, 0.8>, rgb <0.25, 0.3, 0.25> } finish { ambient 0.1 diffuse 0.7 reflection 0.2 } }
sphere { <-3.0, 1.2, 1.0>, 1.2 pigment { rgb <0.9, 0.9, 0.95> } finish { ambient 0.02 diffuse 0.1 reflection 0.85 specular 0.5 roughness 0.01 } }
sphere { <0.5, 1.0, -2.0>, 1.0 pigment { rgbf <0.95, 1, 0.95, 0.9> } finish { ambient 0.02 diffuse 0.1 specular 0.5 roughness 0.01 reflection 0.1 } interior { ior 1.4 } }
box { <2.5, 0, 0.5>, <4.5, 2.4, 2.5> pigment { rgb <0.8, 0.3, 0.2> } finish { ambient 0.1 diffuse 0.7 phong 0.4 } rotate y*20 }
cylinder { <-6, 0, 6>, <-6, 5, 6>, 0.7 pigment { rgb <0.3, 0.5, 0.8> } finish { ambient 0.1 diffuse 0.7 } }
sphere { <7, 2, 12>, 2 hollow pigment { rgbt <0.9, 0.5, 0.3, 0.6> } finish { ambient 0.1 diffuse 0.6 } }
