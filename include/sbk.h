/* sbk.h -- C ABI of the B200 batched forward-dynamics engine ("sbk" = Simbody batched kernels).
 *
 * This is the drop-in boundary for ONE hot path of simbody/simbody 3.9.0: the O(n)
 * articulated-body forward dynamics of SimbodyMatterSubsystem plus the
 * RungeKuttaMersonIntegrator step, evaluated for N independent instances of one
 * tree-topology system at once on one B200 (sm_100a), FP64 throughout.
 *
 * The reference exposes no FFI seam for this path; the boundary is the
 * public operator set of SimbodyMatterSubsystem / Integrator.  Every entry point below
 * names the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain C, opaque handles, int status return (0 = SBK_OK), sbk_last_error() text;
 *     no exceptions cross the boundary, no torch / CUDA types in any signature
 *     (a CUDA stream is passed as void*).
 *   - one sbk_batch is bound to one CUDA device + one stream; not thread-safe per handle.
 *   - the caller owns host buffers, the library owns device buffers.
 *   - batched vectors are structure-of-arrays, slot-major: element (slot i, instance k)
 *     lives at buf[i*N + k]  (SBK_LAYOUT_SOA).  Host entry points also accept
 *     instance-major buffers buf[k*len + i] via the *_aos variants, matching an array
 *     of reference `State` objects.
 *   - q uses the reference's slot rules: slots are handed out in MobilizedBodyIndex
 *     order with max-nq per mobilizer (Pin 1, Slider 1, Universal 2, Ball 4, Free 7;
 *     Simbody/src/RigidBodyNodeSpec.h:81-87), quaternion scalar-first.
 *   - spatial vectors are (angular[3], linear[3]) expressed in Ground, taken at the
 *     body origin (SimTKcommon/Mechanics/include/SimTKcommon/internal/SpatialAlgebra.h:62-108).
 *   - there is NO CPU fallback: every compute entry point fails with SBK_ERR_CUDA when
 *     no sm_100-class device is usable.
 */
#ifndef SBK_H_
#define SBK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBK_VERSION 100  /* 0.1.0 */

/* ---- status codes ------------------------------------------------------------------ */
enum {
    SBK_OK            = 0,
    SBK_ERR_ARG       = 1,  /* wrong length / null pointer; mirrors SimTK_APIARGCHECK
                               (Simbody/src/SimbodyMatterSubsystem.cpp:158-167)          */
    SBK_ERR_STAGE     = 2,  /* operator called before the stage it needs; mirrors
                               SimTK_STAGECHECK (SimbodyMatterSubsystem.h:2138-2139)     */
    SBK_ERR_TOPOLOGY  = 3,  /* not a tree / unsupported mobilizer                        */
    SBK_ERR_CUDA      = 4,  /* CUDA runtime failure or no usable device                  */
    SBK_ERR_NUMERIC   = 5   /* some instance went non-finite / singular D (see status)   */
};

/* ---- mobilizer kinds (Simbody/src/RigidBodyNodeSpec_{Pin,Slider,Universal,Ball,Free}.h) */
enum {
    SBK_JOINT_GROUND    = 0,
    SBK_JOINT_PIN       = 1,  /* nq=1 nu=1, rotation about z of F/M                      */
    SBK_JOINT_SLIDER    = 2,  /* nq=1 nu=1, translation along x of F/M                   */
    SBK_JOINT_UNIVERSAL = 3,  /* nq=2 nu=2, body-fixed x then y                          */
    SBK_JOINT_BALL      = 4,  /* nq=4 nu=3, quaternion (scalar first)                    */
    SBK_JOINT_FREE      = 5,  /* nq=7 nu=6, quaternion + translation in F                */
    SBK_JOINT_WELD      = 6,  /* nq=0 nu=0, F and M coincide (RigidBodyNode_Weld.cpp:369)  */
    SBK_JOINT_TRANSLATION = 7, /* nq=3 nu=3, Cartesian translation in F (RigidBodyNodeSpec_Translation.h) */
    SBK_JOINT_CYLINDER  = 8,  /* nq=2 nu=2, rotation about then translation along z (RigidBodyNodeSpec_Cylinder.h) */
    SBK_JOINT_PLANAR    = 9,  /* nq=3 nu=3, rotation about z, translation x,y in F (RigidBodyNodeSpec_Planar.h)   */
    SBK_JOINT_GIMBAL    = 10  /* nq=3 nu=3, body-fixed x-y-z Euler angles, u = qdot (RigidBodyNodeSpec_Gimbal.h)  */
};

/* ---- force elements (Simbody/src/Force_Gravity.cpp:514-577, Force.cpp:339-351,434-443) */
enum {
    SBK_FORCE_GRAVITY = 1,  /* Force::Gravity: a = magnitude g, dir = unit "down"        */
    SBK_FORCE_SPRING  = 2,  /* Force::MobilityLinearSpring: body, coord, a = k, b = q0   */
    SBK_FORCE_DAMPER  = 3,  /* Force::MobilityLinearDamper: body, coord, a = c           */
    SBK_FORCE_UNIFORM_GRAVITY = 4, /* Force::UniformGravity (Force.cpp:1034-1057): dir = the gravity VECTOR g in Ground, zero height 0 */
    SBK_FORCE_GLOBAL_DAMPER   = 5, /* Force::GlobalDamper (Force.cpp:996-998): f -= a*u on every mobility          */
    SBK_FORCE_MOBILITY_CONSTANT = 6, /* Force::MobilityConstantForce (Force_MobilityConstantForce.h:45): body, coord, a = f */
    SBK_FORCE_TWO_POINT_SPRING = 7, /* Force::TwoPointLinearSpring (Force.cpp:103-140): body = body1, coord = body2, a = k, b = x0,
                                       dir = station1 (in B1), station2 (in B2); body 0 is Ground */
    SBK_FORCE_TWO_POINT_DAMPER = 8  /* Force::TwoPointLinearDamper (Force.cpp:179-221): body = body1, coord = body2, a = damping,
                                       dir = station1, station2 */
};

/* One mobilized body, in MobilizedBodyIndex order; entry 0 must be Ground.
 * Filled from a realized MultibodySystem by simbody_b200/host/lower_simbody.h using
 * MobilizedBody::getParentMobilizedBody / getDefaultInboardFrame / getDefaultOutboardFrame /
 * getDefaultMassProperties (Simbody/include/simbody/internal/MobilizedBody.h:1556-1667). */
typedef struct sbk_body_desc {
    int32_t parent;               /* MobilizedBodyIndex of the parent; -1 for Ground     */
    int32_t joint_type;           /* SBK_JOINT_*                                          */
    double  mass;
    double  com_B[3];             /* mass centre, from Bo, in B                           */
    double  unit_inertia_OB_B[6]; /* xx yy zz xy xz yz, about Bo, in B, per unit mass     */
    double  X_PF[12];             /* R_PF row-major (9) then p_PF (3)                     */
    double  X_BM[12];             /* R_BM row-major (9) then p_BM (3)                     */
} sbk_body_desc;

typedef struct sbk_force_desc {
    int32_t kind;                 /* SBK_FORCE_*                                          */
    int32_t body;                 /* spring/damper: MobilizedBodyIndex                    */
    int32_t coord;                /* spring: MobilizerQIndex; damper: MobilizerUIndex     */
    int32_t pad_;
    double  a;                    /* gravity g | spring k | damper c | global damper c    */
    double  b;                    /* spring q0                                            */
    double  dir[3];               /* gravity: unit down direction | uniform gravity: vector g | two-point: station on body1 */
    double  station2[3];          /* two-point elements: station on body2 (in B2)           */
} sbk_force_desc;

/* Integrator options; mirrors Integrator::setAccuracy / setConstraintTolerance /
 * setUseInfinityNorm / setProjectEveryStep (SimTKmath/Integrators/include/simmath/Integrator.h:352-394).
 * Fixed-step only in this release (setFixedStepSize + setAllowInterpolation(false)).  */
typedef struct sbk_rkm_opts {
    double  accuracy;             /* default 1e-3 (IntegratorRep.h:737)                   */
    double  constraint_tol;       /* default accuracy/10 (IntegratorRep.h:738-740)        */
    int32_t use_infinity_norm;    /* 0 = weighted RMS (default), 1 = Inf norm             */
    int32_t project_every_step;   /* 0 default; 1 = always normalise quaternions          */
} sbk_rkm_opts;

/* Error-controlled stepping; mirrors Integrator::setAccuracy / setInitialStepSize /
 * setMinimumStepSize / setMaximumStepSize / setAllowInterpolation
 * (SimTKmath/Integrators/include/simmath/Integrator.h:352-394).  Values <= 0 mean "default":
 * accuracy 1e-3, constraint_tol accuracy/10, init_step timescale/10 = 0.01
 * (AbstractIntegratorRep.cpp:40-53), no min/max step.                                       */
typedef struct sbk_adaptive_opts {
    double  accuracy, constraint_tol, init_step, min_step, max_step;
    int32_t use_infinity_norm, project_every_step;
    int32_t allow_interpolation;   /* 0 (default) = TimeStepper::stepTo semantics (README example):
                                      t_final bounds the internal steps, the last one lands on it
                                      (hWasArtificiallyLimited logic, AbstractIntegratorRep.cpp:533-541);
                                      1 = bare Integrator::stepTo(reportTime) with interpolation
                                      allowed: steps are never shortened, the advanced state ends at
                                      t >= t_final (the reference then reports an interpolated state) */
    int32_t max_attempts;          /* safety cap on step attempts per instance per call (0 = 1e6) */
} sbk_adaptive_opts;

typedef struct sbk_topology sbk_topology;
typedef struct sbk_batch    sbk_batch;

/* ---- library ------------------------------------------------------------------------ */
int         sbk_version(void);
const char* sbk_last_error(void);
/* Number of usable sm_100-class CUDA devices (0 when none; never falls back to CPU).    */
int         sbk_device_count(void);

/* ---- topology: replaces SimbodyMatterSubsystemRep::endConstruction
 *      (Simbody/src/SimbodyMatterSubsystemRep.cpp:256-330) for the supported subset ------ */
/* Flags of sbk_topology_create_ex.  SBK_TOPOLOGY_EULER_ANGLES mirrors SimbodyMatterSubsystem::setUseEulerAngles
 * (SimbodyMatterSubsystem.h): Ball / Free orientations are three body-fixed x-y-z angles instead of a
 * quaternion (RigidBodyNodeSpec_Ball.h:118-160, _Free.h:147-250); the q slots keep their quaternion-sized
 * allocation (the last one is unused), there are no quaternion constraints to project.            */
enum { SBK_TOPOLOGY_EULER_ANGLES = 1 };
sbk_topology* sbk_topology_create_ex(const sbk_body_desc* bodies, int nb,
                                     const sbk_force_desc* forces, int nf, unsigned flags);
sbk_topology* sbk_topology_create(const sbk_body_desc* bodies, int nb,
                                  const sbk_force_desc* forces, int nf);
void sbk_topology_destroy(sbk_topology*);
/* counts: nb (incl. Ground), nq, nu, number of quaternions, number of levels (incl. level 0) */
int  sbk_topology_counts(const sbk_topology*, int* nb, int* nq, int* nu, int* nquat, int* nlevels);
/* per-body slot map (arrays of length nb): first q, nq, first u, nu, level              */
int  sbk_topology_slots(const sbk_topology*, int* q0, int* nq, int* u0, int* nu, int* level);
/* Parse / build model descriptions (text format of simbody_b200/host/model_spec.h).
 * sbk_model_text writes the text of a named built-in model ("double_pendulum",
 * "pin_chain", "humanoid30", "branched_tree", "mixed7") with size parameter n into buf;
 * returns the length needed (incl. NUL) or -1.                                          */
int  sbk_model_text(const char* name, int n, char* buf, int cap);
sbk_topology* sbk_topology_from_text(const char* model_text);

/* ---- batch -------------------------------------------------------------------------- */
/* stream: a cudaStream_t passed as void* (NULL = the library creates its own).          */
sbk_batch* sbk_batch_create(const sbk_topology*, int n_instances, int device, void* stream);
void sbk_batch_destroy(sbk_batch*);
int  sbk_batch_size(const sbk_batch*);
/* Execution plan: 0 = auto, 1 = generic thread-per-instance, 2 = register-resident fused
 * (small models), 3 = level-parallel CTA-per-instance (wide trees), 4 = grid-level-parallel
 * integrator (wide trees, small batches: the whole GPU works on one tree level of the batch;
 * other operations run as in plan 1).                                                    */
int  sbk_batch_set_plan(sbk_batch*, int plan);
int  sbk_batch_get_plan(const sbk_batch*);
int  sbk_synchronize(sbk_batch*);

/* State access; replaces State::updQ/updU/getQ/getU/getTime
 * (SimTKcommon/Simulation/include/SimTKcommon/internal/State.h:962-1043).
 * Host buffers: q [nq][N], u [nu][N] (SoA), t [N] or NULL.  Setting state invalidates
 * all realized stages, as in the reference.                                             */
int sbk_set_state(sbk_batch*, const double* q, const double* u, const double* t);
int sbk_get_state(sbk_batch*, double* q, double* u, double* t);
/* Asynchronous variants for PINNED (page-locked) host buffers: the copies are queued on the batch's stream and the call
 * returns at once; the buffers must stay untouched until sbk_synchronize().  (State::updQ()/getQ() have no asynchronous
 * analogue in the reference; these exist so that a set -> step -> get round trip costs one synchronisation.) */
int sbk_set_state_async(sbk_batch*, const double* q, const double* u, const double* t);
int sbk_get_state_async(sbk_batch*, double* q, double* u, double* t);
int sbk_set_state_aos(sbk_batch*, const double* q /*[N][nq]*/, const double* u /*[N][nu]*/);
int sbk_get_state_aos(sbk_batch*, double* q, double* u);
/* Device pointers of the resident SoA state (for zero-copy plumbing, e.g. torch).       */
int sbk_state_device_ptrs(sbk_batch*, double** q, double** u, double** t);
/* Tell the library the resident state was modified through those pointers.              */
int sbk_state_touched(sbk_batch*);

/* ---- realize stages ------------------------------------------------------------------ */
/* SimbodyMatterSubsystem::realizePositionKinematics (SimbodyMatterSubsystem.h:2691)      */
int sbk_realize_position(sbk_batch*);
/* ...::realizeVelocityKinematics (SimbodyMatterSubsystem.h:2705); implies Position       */
int sbk_realize_velocity(sbk_batch*);
/* ...::realizeArticulatedBodyInertias (SimbodyMatterSubsystem.h:2730); implies Position  */
int sbk_realize_articulated_body_inertias(sbk_batch*);
/* System::realize(state, Stage::Acceleration) for the lowered force set
 * (Simbody/src/MultibodySystem.cpp:248-333): kinematics, gravity/spring/damper forces,
 * ABI, both acceleration sweeps, qdot/qdotdot.                                          */
int sbk_realize_acceleration(sbk_batch*);

/* ---- results of the last realize (host, SoA) ------------------------------------------ */
int sbk_get_udot(sbk_batch*, double* udot /*[nu][N]*/);
int sbk_get_qdot(sbk_batch*, double* qdot /*[nq][N]*/);
int sbk_get_qdotdot(sbk_batch*, double* qdotdot /*[nq][N]*/);
int sbk_get_qerr(sbk_batch*, double* qerr /*[nquat][N]*/);
/* MobilizedBody::getBodyTransform / getBodyVelocity / getBodyAcceleration
 * (MobilizedBody.h:316,348,377).  X: [nb][12][N] (R row-major, p); V, A: [nb][6][N].     */
int sbk_get_body_transforms(sbk_batch*, double* X_GB);
int sbk_get_body_velocities(sbk_batch*, double* V_GB);
int sbk_get_body_accelerations(sbk_batch*, double* A_GB);
/* Applied forces accumulated by the last sbk_realize_acceleration:
 * mobility forces [nu][N], body forces [nb][6][N] (MultibodySystem.cpp:163-177).         */
int sbk_get_applied_forces(sbk_batch*, double* f_mob, double* F_body);

/* MultibodySystem::calcKineticEnergy / calcPotentialEnergy for the lowered system
 * (SimbodyMatterSubsystemRep.cpp:5234-5246, RigidBodyNode.cpp:182-188, Force_Gravity.cpp:555,
 * Force.cpp:354-361).  Needs Velocity stage (ke) / Position stage (pe).  Host [N], nullable.   */
int sbk_calc_energy(sbk_batch*, double* kinetic, double* potential);

/* SimbodyMatterSubsystem::calcMobilizerReactionForces (SimbodyMatterSubsystem.h:2479,
 * SimbodyMatterSubsystemRep.cpp:5788-5832): the spatial force each mobilizer applies to its outboard
 * body at the origin of the outboard frame M, expressed in Ground; entry 0 is the reaction that
 * holds Ground.  Needs the acceleration stage (SBK_ERR_STAGE otherwise, also after an operator call
 * that reused the acceleration cache).  Host FM_G [nb][6][N].                                  */
int sbk_calc_mobilizer_reaction_forces(sbk_batch*, double* FM_G);

/* SimbodyMatterSubsystem::multiplyBySystemJacobian / multiplyBySystemJacobianTranspose
 * (SimbodyMatterSubsystem.h:554,646; RigidBodyNodeSpec.cpp:760-815).  Position stage.
 * v [nu][N] -> Jv [nb][6][N] (body origin spatial velocities);  F [nb][6][N] -> JtF [nu][N].  */
int sbk_multiply_by_system_jacobian(sbk_batch*, const double* v, double* Jv);
int sbk_multiply_by_system_jacobian_transpose(sbk_batch*, const double* F_body, double* JtF);

/* SimbodyMatterSubsystem::calcCompositeBodyInertias (SimbodyMatterSubsystemRep.cpp:5196-5205,
 * RigidBodyNode.cpp:231-243): for every body the spatial inertia of the rigid body obtained by
 * locking all joints outboard of it, about the body origin, in Ground.  Position stage.
 * Host R [nb][10][N]: mass, mass centre (3), unit inertia xx yy zz xy xz yz; Ground = infinite mass. */
int sbk_calc_composite_body_inertias(sbk_batch*, double* R);

/* ---- operators ------------------------------------------------------------------------ */
/* SimbodyMatterSubsystem::calcAcceleration / calcAccelerationIgnoringConstraints
 * (SimbodyMatterSubsystem.h:2141,2171; SimbodyMatterSubsystem.cpp:151-226).
 * Requires velocity stage.  f_mob [nu][N] and F_body [nb][6][N] may each be NULL (= zero,
 * like the reference's zero-length vectors).  udot [nu][N], A_GB [nb][6][N] (either may
 * be NULL).                                                                             */
int sbk_calc_acceleration(sbk_batch*, const double* f_mob, const double* F_body,
                          double* udot, double* A_GB);
/* SimbodyMatterSubsystem::multiplyByM (SimbodyMatterSubsystem.h:1262); needs Position.   */
int sbk_multiply_by_M(sbk_batch*, const double* a /*[nu][N]*/, double* Ma /*[nu][N]*/);
/* SimbodyMatterSubsystem::multiplyByMInv (SimbodyMatterSubsystem.h:1343); needs Position,
 * realizes ABI lazily like the reference (SimbodyMatterSubsystemRep.cpp:5364).           */
int sbk_multiply_by_MInv(sbk_batch*, const double* v /*[nu][N]*/, double* MinvV /*[nu][N]*/);
/* SimbodyMatterSubsystem::calcResidualForceIgnoringConstraints (SimbodyMatterSubsystem.h:2234);
 * needs Velocity.  f_mob, F_body, known_udot may be NULL (= zero, SimbodyMatterSubsystemRep.cpp:5546-5564). */
int sbk_calc_residual_force(sbk_batch*, const double* f_mob, const double* F_body,
                            const double* known_udot, double* residual /*[nu][N]*/);

/* ---- integrator ----------------------------------------------------------------------- */
void sbk_rkm_default_opts(sbk_rkm_opts*);
/* RungeKuttaMersonIntegrator with setFixedStepSize(h), setAllowInterpolation(false):
 * nsteps accepted steps of size h for every instance
 * (SimTKmath/Integrators/src/RungeKuttaMersonIntegrator.cpp:86-140,
 *  AbstractIntegratorRep.cpp:137-208,513-578).  The state stays resident on the device.
 * err_norm (host, [N], nullable) receives the error norm of the LAST step
 * (IntegratorRep.h:454-488).                                                            */
int sbk_rkm_step(sbk_batch*, double h, int nsteps, const sbk_rkm_opts* opts, double* err_norm);
/* RungeKuttaMersonIntegrator with error control: Integrator::stepTo(t_final) for every instance,
 * each with its own step size history (AbstractIntegratorRep.cpp:216-368 stepping loop,
 * :448-502 adjustStepSize, :513-578 takeOneStep).  Per-instance outputs (host, [N], nullable):
 * internal steps taken and attempted since the state was last set, and the last accepted step
 * size.  sbk_get_state's t returns each instance's advanced time.  Not available in plans 3 and 4. */
void sbk_adaptive_default_opts(sbk_adaptive_opts*);
int sbk_rkm_adaptive(sbk_batch*, double t_final, const sbk_adaptive_opts* opts,
                     int32_t* steps_taken, int32_t* steps_attempted, double* last_step);
/* Integrator::getNumStepsTaken / getNumRealizations / getNumQProjections
 * (Integrator.h:286-290): totals over the batch since creation.                         */
int sbk_rkm_stats(sbk_batch*, int64_t* steps_taken, int64_t* realizations, int64_t* q_projections);
/* Per-instance status word, OR-accumulated over the operations since the last sbk_set_state* (which clears it):
 *   bit 0 (1) non-finite error norm of an integrator step (NaN / Inf state, or a step whose constraint violation exceeded the
 *             projection limit, AbstractIntegratorRep.cpp:165-190)
 *   bit 1 (2) singular joint-space inertia D = ~H P H (RigidBodyNodeSpec.cpp:293 would divide by it)
 *   bit 2 (4) sbk_rkm_adaptive ran out of its attempt budget before reaching t_final
 * n_bad = number of instances with a non-zero word. */
int sbk_get_status(sbk_batch*, int32_t* status /*[N]*/, int64_t* n_bad);
/* Number of kernels launched by this batch since creation (for bench accounting).       */
int64_t sbk_launch_count(const sbk_batch*);
/* Name of the fixed-step integrator kernel the batch's plan launches, spelled like a profiler's demangled name
 * (e.g. "tpiKernel<7, 1, 2, 1073741886, 1>"): lets a bench line be matched with a committed ncu capture. */
int sbk_integrator_kernel_name(const sbk_batch*, char* buf, int cap);
/* Device time in ms of the kernels launched by the last sbk_rkm_step call, measured with
 * CUDA events on the batch's stream (0 if events disabled).                              */
double  sbk_last_kernel_ms(const sbk_batch*);

/* FP64 roofline probe used by bench.py: runs blocks*threads threads each doing iters*8
 * dependent-chain-free DFMAs on `device`, returns the kernel time in ms (CUDA events).
 * flops = 2 * 8 * iters * blocks * threads.                                               */
int sbk_dfma_probe(int device, int blocks, int threads, int iters, double* ms_out);

/* Diagnostics: time `sweeps` passes of the thread-per-instance record access pattern (nb records
 * of rows_in loaded + rows_out stored doubles per instance, [row][N] layout) with no arithmetic;
 * bytes moved = 8*n*nb*(rows_in+rows_out)*sweeps.  min_blocks 2 or 4 selects the occupancy.     */
int sbk_mem_pattern_probe(int device, int n, int nb, int rows_in, int rows_out, int sweeps, int min_blocks, double* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* SBK_H_ */
