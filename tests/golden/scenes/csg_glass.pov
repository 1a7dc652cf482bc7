// CSG + refraction + filtered shadows (config-3 flavour, small)
#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 6 }
background { rgb <0.05, 0.07, 0.12> }
camera { location <0.3, 3.2, -7.5> look_at <0, 0.9, 0> angle 42 right x*16/9 }
light_source { <6, 9, -7> rgb <1, 0.95, 0.9> }
light_source { <-8, 5, -3> rgb <0.3, 0.35, 0.5> }
plane { y, -0.0078125 pigment { checker rgb <0.9,0.9,0.9>, rgb <0.2,0.25,0.3> scale 0.75 } finish { ambient 0.1 diffuse 0.7 } }
difference {
  box { <-1, 0, -1>, <1, 2, 1> }
  sphere { <0, 1, 0>, 1.25 }
  pigment { rgbf <0.7, 0.9, 1.0, 0.7> } finish { ambient 0.05 diffuse 0.2 specular 0.6 roughness 0.01 reflection 0.1 }
  interior { ior 1.45 }
  translate <-2.1, 0.01, 0.4>
}
intersection {
  sphere { <0, 1, 0>, 1.1 }
  box { <-0.8, 0.1, -0.8>, <0.8, 1.9, 0.8> rotate y*30 }
  quadric { <1, 0, 1>, <0, 0, 0>, <0, 0, 0>, -0.55 }
  pigment { rgb <0.9, 0.5, 0.2> } finish { ambient 0.1 diffuse 0.6 phong 0.6 phong_size 60 }
  translate <0.4, 0.01, 1.2>
}
merge {
  sphere { <0, 0.7, 0>, 0.7 }
  sphere { <0.6, 1.0, 0>, 0.5 }
  box { <-0.4, 0, -0.4>, <0.4, 1.5, 0.4> }
  pigment { rgbt <0.5, 1.0, 0.6, 0.6> } finish { ambient 0.1 diffuse 0.4 specular 0.3 }
  interior { ior 1.3 fade_distance 1.5 fade_power 2 fade_color <0.2, 0.9, 0.4> }
  translate <2.3, 0.01, -0.3>
}
difference {
  sphere { 0, 1 scale <1.3, 0.6, 0.9> }
  quadric { <1, 1, 0>, <0, 0, 0>, <0, 0, 0>, -0.16 }
  plane { y, 0 rotate z*20 }
  pigment { rgb <0.8, 0.2, 0.3> } finish { ambient 0.1 diffuse 0.6 reflection 0.25 }
  translate <0.2, 0.9, -1.9>
}
