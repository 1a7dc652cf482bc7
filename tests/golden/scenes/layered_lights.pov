// layered textures, transformed boxes, ellipsoids, cylinder / spot / fill lights, metallic + fresnel finishes, no_shadow / no_image flags
#version 3.7;
global_settings { assumed_gamma 1 ambient_light rgb <0.8, 0.8, 1.0> }
background { rgb <0.02, 0.02, 0.03> }
camera { location <0, 3, -9> look_at <0, 1, 0> angle 40 right x*16/9 }
light_source { <4, 8, -6> rgb <1,1,1> fade_distance 8 fade_power 2 }
light_source { <-5, 7, -2> rgb <0.8, 0.6, 0.3> cylinder point_at <-1, 0, 0> radius 2 falloff 3 tightness 1 }
light_source { <0, 2, -12> rgb 0.2 shadowless }
plane { y, -0.0078125 pigment { rgb <0.6, 0.6, 0.65> } finish { ambient 0.1 diffuse 0.6 reflection { 0.05, 0.4 falloff 2 } } }
box { <-0.7, 0, -0.7>, <0.7, 1.4, 0.7>
  texture { pigment { checker rgb <0.9,0.1,0.1>, rgb <0.9,0.9,0.1> scale 0.35 } finish { ambient 0.1 diffuse 0.7 } }
  texture { pigment { gradient y color_map { [0 rgbt <0,0,1,0.2>] [1 rgbt <0,0,1,1>] } scale 1.5 } finish { phong 0.8 } }
  rotate <0, 35, 10> translate <-2.4, 0.3, 0.5> }
sphere { 0, 1 scale <1.2, 0.7, 0.8> rotate z*25 translate <0.2, 1.0, 0.3>
  pigment { rgb <0.9, 0.75, 0.3> } finish { ambient 0.05 diffuse 0.4 specular 0.8 roughness 0.02 metallic reflection { 0.5 metallic } } }
sphere { <2.6, 0.9, 0.6>, 0.9 pigment { rgb <0.2, 0.5, 0.9> } finish { ambient 0.1 diffuse 0.5 brilliance 2.5 phong 0.4 phong_size 80 } }
sphere { <1.3, 0.4, -2.0>, 0.4 pigment { rgb <0.3, 0.9, 0.4> } finish { diffuse 0.6 } no_shadow }
sphere { <-1.0, 0.45, -2.4>, 0.45 pigment { rgb <0.9, 0.3, 0.8> } finish { diffuse 0.6 } no_image }
box { <-0.5, 0, -0.5>, <0.5, 0.6, 0.5> pigment { rgb <0.7,0.7,0.7> } finish { diffuse 0.3 reflection 0.6 } no_reflection translate <3.0, 0.0, -2.2> }
box { <-3, 0, 3>, <3, 3, 3.2> pigment { rgb <0.4, 0.45, 0.5> } finish { diffuse 0.5 reflection 0.3 conserve_energy } }
