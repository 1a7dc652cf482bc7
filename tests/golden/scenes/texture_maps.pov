// texture_map (blends whole textures: different finishes, normals), nested maps, average texture_map,
// block-pattern texture lists, transparent entries in shadows
#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 4 }
camera { location <0, 5, -11> look_at <0, 1.0, 0> angle 46 right x*16/9 }
light_source { <12, 18, -14> rgb <1, 1, 1> }
light_source { <-8, 6, -6> rgb <0.3, 0.3, 0.35> }
background { rgb <0.06, 0.08, 0.12> }
#declare T_Shiny = texture { pigment { rgb <0.9, 0.2, 0.2> } finish { ambient 0.1 diffuse 0.6 phong 0.8 reflection 0.25 } }
#declare T_Matte = texture { pigment { bozo color_map { [0 rgb <0.2, 0.5, 0.2>] [1 rgb <0.8, 0.9, 0.4>] } scale 0.2 } normal { bumps 0.5 scale 0.1 } finish { ambient 0.15 diffuse 0.8 } }
#declare T_Layered = texture { pigment { gradient x color_map { [0.3 rgb <0.2, 0.3, 0.8>] [0.7 rgbt <1, 0.9, 0.2, 0.1>] } scale 0.3 } normal { dents 0.6 scale 0.2 } finish { ambient 0.1 diffuse 0.6 specular 0.5 } }
plane { y, 0 texture { checker texture { T_Matte } texture { pigment { rgb <0.85, 0.85, 0.9> } finish { ambient 0.1 diffuse 0.7 reflection 0.2 } } scale 1.5 } }
sphere { <-4.0, 1.2, 0.5>, 1.2 texture { gradient y texture_map { [0.2 T_Shiny] [0.5 T_Matte] [0.9 T_Layered] } scale 2.4 translate -0.05*y } }
sphere { <-1.3, 1.2, 0.5>, 1.2 texture { bozo turbulence 0.3 texture_map { [0.35 T_Shiny] [0.65 marble texture_map { [0 T_Matte] [1 T_Layered] } scale 0.4 rotate z*40] } scale 0.6 } }
sphere { <1.4, 1.2, 0.5>, 1.2 texture { average texture_map { [1 T_Shiny] [2 T_Layered] [1 pigment { rgb <1, 1, 1> } finish { ambient 0.3 diffuse 0.5 }] } } }
sphere { <4.1, 1.2, 0.5>, 1.2 texture { wrinkles texture_map { [0.3 pigment { rgbf <0.9, 1, 0.9, 0.8> } finish { ambient 0.02 diffuse 0.2 specular 0.4 }] [0.7 T_Shiny] } scale 0.5 } interior { ior 1.3 } }
box { <-1.5, 0.05, -3.8>, <1.5, 1.0, -2.6> texture { hexagon texture { T_Shiny } texture { T_Matte } texture { T_Layered } scale 0.4 rotate x*90 } }
