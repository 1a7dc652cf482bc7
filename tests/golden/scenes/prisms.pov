// prisms (prism.cpp:194-770): linear / quadratic / cubic / bezier splines, linear and conic sweeps, open and closed, sturm,
// as a CSG operand and transparent (spline normals under refraction)
#version 3.7;
global_settings { assumed_gamma 1.0 max_trace_level 6 }
camera { location <0, 6.5, -11> look_at <0, 0.6, 0> angle 44 }
light_source { <-8, 12, -9> rgb <1, 0.95, 0.9> }
light_source { <9, 6, -4> rgb <0.35, 0.4, 0.55> }
plane { y, -0.01 pigment { checker rgb <0.85, 0.85, 0.8>, rgb <0.3, 0.35, 0.45> scale 1.5 } }
prism { linear_sweep linear_spline 0, 1, 7, <3,5>, <-3,5>, <-5,0>, <-3,-5>, <3,-5>, <5,0>, <3,5>
        pigment { rgb <0.9, 0.5, 0.2> } finish { phong 0.5 } scale <0.25, 1, 0.25> translate <-4.2, 0, 2.5> }
prism { linear_sweep quadratic_spline 0, 1.2, 8, <-1,-1>, <0,-2>, <2,-1>, <2.5,1>, <0,2>, <-2,1>, <0,-2>, <2,-1>
        open pigment { rgb <0.3, 0.7, 0.4> } finish { specular 0.4 } scale <0.6, 1, 0.6> translate <-1.3, 0, 2.8> }
prism { linear_sweep cubic_spline 0, 0.9, 18,
        <3,-5>, <3,5>, <-5,0>, <3,-5>, <3,5>, <-5,0>,          // outer triangle (control points wrap)
        <2,-4>, <2,4>, <-4,0>, <2,-4>, <2,4>, <-4,0>,          // inner triangle: a hole
        <1,-1.5>, <1,1.5>, <-1.5,0>, <1,-1.5>, <1,1.5>, <-1.5,0>
        sturm pigment { rgb <0.5, 0.55, 0.9> } scale <0.3, 1, 0.3> translate <1.8, 0, 2.6> }
prism { linear_sweep bezier_spline 0, 1.1, 8,
        <0,-2>, <2.5,-2>, <2.5,2>, <0,2>,  <0,2>, <-2.5,2>, <-2.5,-2>, <0,-2>
        pigment { rgbf <0.9, 1, 0.95, 0.75> } finish { specular 0.5 reflection 0.1 } interior { ior 1.4 }
        scale <0.45, 1, 0.45> translate <4.4, 0, 2.2> }
prism { conic_sweep linear_spline 0, 1, 5, <4,4>, <-4,4>, <-4,-4>, <4,-4>, <4,4>
        pigment { rgb <0.9, 0.8, 0.3> } rotate 180 * x translate y scale <0.3, 1.6, 0.3> translate <-3.4, 0, -1.2> }
prism { conic_sweep cubic_spline 0.3, 1, 7, <0,3>, <3,0>, <0,-3>, <-3,0>, <0,3>, <3,0>, <0,-3>
        pigment { rgb <0.85, 0.3, 0.35> } finish { phong 0.6 } rotate 180 * x translate y scale <0.5, 1.8, 0.5> translate <-0.6, 0, -1.6> }
difference {
  prism { linear_sweep cubic_spline -0.2, 1.0, 7, <-2,0>, <0,-2>, <2,0>, <0,2>, <-2,0>, <0,-2>, <2,0> }
  cylinder { <0, -1, 0>, <0, 2, 0>, 0.7 }
  box { <0, 0.5, -3>, <3, 2, 0> }
  pigment { rgb <0.4, 0.8, 0.85> } finish { specular 0.3 }
  scale 0.75 rotate 25 * y translate <2.9, 0.2, -1.5>
}
