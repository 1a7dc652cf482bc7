#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 5 }
background { rgb <0.08, 0.1, 0.16> }
camera { perspective location <0, 5, -12> direction <0, 0, 1.7> up <0, 1, 0> right <1.7777777777777777, 0, 0> look_at <0, 1.2, 0> }
light_source { <10, 16, -12> rgb <1, 1, 1> }
light_source { <-9, 7, -5> rgb <0.3, 0.35, 0.45> }
plane { y, -0.0078125 pigment { checker rgb <0.85, 0.85, 0.85>, rgb <0.25, 0.3, 0.35> } finish { ambient 0.1 diffuse 0.7 } }
cylinder { <-4, 0, 0>, <-4, 2.5, 0.5>, 0.7 pigment { rgb <0.9, 0.3, 0.2> } finish { ambient 0.1 diffuse 0.6 phong 0.5 } }
cylinder { <-2.2, 0.4, 1>, <-0.8, 1.2, -0.5>, 0.45 open pigment { rgb <0.2, 0.7, 0.9> } finish { ambient 0.1 diffuse 0.6 specular 0.4 } }
cone { <0.5, 0, 0>, 1.1, <0.5, 2.4, 0>, 0.2 pigment { rgb <0.9, 0.8, 0.2> } finish { ambient 0.1 diffuse 0.65 reflection 0.15 } }
cone { <2.6, 0.2, 1.0>, 0.0, <2.2, 2.0, 0.2>, 0.9 open pigment { rgb <0.6, 0.3, 0.8> } finish { ambient 0.1 diffuse 0.6 } }
cone { <4.4, 0, -0.5>, 0.9, <4.4, 1.8, -0.5>, 0.9001 scale <1, 1, 0.6> rotate <0, 20, 0> pigment { rgb <0.4, 0.9, 0.4> } finish { ambient 0.1 diffuse 0.6 } }
// CSG over cylinders / cones: a drilled block, a pipe, a glass cone with a cylindrical hole
difference { box { <-1.2, 0, -3.6>, <1.2, 1.4, -1.8> } cylinder { <0, 0.7, -4>, <0, 0.7, -1.4>, 0.45 } cylinder { <-1.5, 0.7, -2.7>, <1.5, 0.7, -2.7>, 0.3 }
  pigment { rgb <0.8, 0.5, 0.3> } finish { ambient 0.1 diffuse 0.6 phong 0.3 } }
difference { cylinder { <3, 0, -3>, <3, 1.6, -3>, 0.8 } cylinder { <3, -0.1, -3>, <3, 1.7, -3>, 0.55 }
  pigment { rgb <0.7, 0.7, 0.75> } finish { ambient 0.1 diffuse 0.5 specular 0.6 roughness 0.02 reflection 0.2 } }
intersection { cone { <-3.2, 0, -3>, 1.0, <-3.2, 2.0, -3>, 0.1 } cylinder { <-3.2, 0.3, -4.5>, <-3.2, 0.3, -1.5>, 0.7 inverse }
  pigment { rgbf <0.85, 0.95, 1.0, 0.7> } finish { ambient 0.02 diffuse 0.3 specular 0.5 roughness 0.02 } interior { ior 1.4 } }
merge { cylinder { <-5.5, 0, -2.5>, <-5.5, 1.2, -2.5>, 0.5 } cone { <-5.5, 1.0, -2.5>, 0.7, <-5.5, 1.9, -2.5>, 0.0 }
  pigment { rgbf <1.0, 0.8, 0.8, 0.6> } finish { ambient 0.02 diffuse 0.4 } interior { ior 1.3 } }
disc { <1.8, 0.6, -5.0>, <0.2, 1, -0.3>, 0.9 pigment { rgb <0.9, 0.6, 0.7> } finish { ambient 0.1 diffuse 0.7 } }
disc { <-1.6, 1.1, -5.2>, <0.5, 0.4, -1>, 0.8, 0.35 pigment { rgb <0.5, 0.8, 0.9> } finish { ambient 0.1 diffuse 0.6 specular 0.3 } }
intersection { sphere { <5.4, 0.9, -3.4>, 0.9 } disc { <5.4, 1.2, -3.4>, <0, 1, 0.2>, 5 }
  pigment { rgb <0.9, 0.9, 0.4> } finish { ambient 0.1 diffuse 0.6 } }
