// pigment_map (nested pigments, evaluated at the parent's warped point), average pigments, block-pattern pigment lists
#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 4 }
camera { location <0, 5, -11> look_at <0, 1.0, 0> angle 46 right x*16/9 }
light_source { <12, 18, -14> rgb <1, 1, 1> }
background { rgb <0.06, 0.08, 0.12> }
plane { y, 0
  pigment { checker pigment { marble turbulence 0.5 color_map { [0 rgb <0.9, 0.9, 0.9>] [1 rgb <0.4, 0.4, 0.5>] } scale 0.4 }
                    pigment { granite color_map { [0 rgb <0.2, 0.3, 0.2>] [1 rgb <0.7, 0.8, 0.6>] } scale 0.6 } scale 1.5 }
  finish { ambient 0.1 diffuse 0.7 } }
sphere { <-3.5, 1.2, 0.5>, 1.2
  pigment { gradient y pigment_map { [0.0 checker rgb <1, 0.2, 0.2>, rgb <1, 1, 1> scale 0.2]
                                     [0.5 bozo color_map { [0 rgb <0.1, 0.2, 0.8>] [1 rgb <0.9, 0.9, 0.3>] } scale 0.15]
                                     [1.0 rgb <0.2, 0.8, 0.3>] } scale 2.4 translate -0.1*y }
  finish { ambient 0.1 diffuse 0.7 phong 0.4 } }
sphere { <0.0, 1.2, 0.5>, 1.2
  pigment { average pigment_map { [1 gradient x color_map { [0 rgb <1, 0, 0>] [1 rgb <0, 0, 1>] } scale 0.5]
                                  [2 onion color_map { [0 rgb <1, 1, 0>] [1 rgb <0, 0.4, 0>] } scale 0.3]
                                  [0.5 rgb <1, 1, 1>] } rotate z*20 }
  finish { ambient 0.1 diffuse 0.7 } }
sphere { <3.5, 1.2, 0.5>, 1.2
  pigment { bozo turbulence 0.3 pigment_map { [0.3 wrinkles color_map { [0 rgb <0.9, 0.5, 0.1>] [1 rgb <0.3, 0.1, 0>] } scale 0.2]
                                              [0.7 gradient y pigment_map { [0 rgb <0.1, 0.6, 0.9>] [1 hexagon rgb <1, 1, 1>, rgb <0.5, 0.5, 0.5>, rgb <0.1, 0.1, 0.1> scale 0.2 rotate x*90] } scale 0.8] } scale 0.7 }
  finish { ambient 0.1 diffuse 0.7 specular 0.3 } }
box { <-1.2, 0.05, -3.8>, <1.2, 0.95, -2.6>
  pigment { average pigment_map { [1 rgbf <1, 0.5, 0.2, 0.3>] [3 rgbf <0.2, 0.5, 1, 0.6>] } } finish { ambient 0.1 diffuse 0.6 } interior { ior 1.2 } }
