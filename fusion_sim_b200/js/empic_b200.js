// empic_b200.js -- drop-in replacement for public/javascripts/empic.js in a headless Node driver
// (SOURCE ONLY: no Node.js in this image; see INTEGRATION.md).  Same export, same ten members
// (empic.js:60, :1157-:1526); nested JS arrays are flattened exactly as set() indexes them.
'use strict';
const native = require('./fusionsim.node');

function validate(spec) {  // utilities.validate_object, utilities.js:118-127
  for (const p of ['radius', 'height', 'nr', 'nz', 'dt', 'nparticles', 'particle_mass', 'particle_charge']) {
    if (typeof spec[p] === 'undefined') throw new Error('.' + p + ' <- Non-optional property is undefined!');
    if (typeof spec[p] !== 'number') throw new Error('.' + p + ' <- Property does not match any given possible types!');
  }
}

function flat3(a, n0, n1) {  // value.E[i][j][k] -> (i*n1 + j)*3 + k
  const out = new Float64Array(n0 * n1 * 3);
  for (let i = 0; i < n0; i++) for (let j = 0; j < n1; j++) for (let k = 0; k < 3; k++) out[(i * n1 + j) * 3 + k] = a[i][j][k];
  return out;
}
function flat2(a) {
  const n0 = a.length, n1 = a[0].length, out = new Float64Array(n0 * n1);
  for (let i = 0; i < n0; i++) for (let j = 0; j < n1; j++) out[i * n1 + j] = a[i][j];
  return out;
}
function flatN(a, w) {
  const out = new Float64Array(a.length * w);
  for (let i = 0; i < a.length; i++) for (let k = 0; k < w; k++) out[i * w + k] = a[i][k];
  return out;
}

exports.makeCylindricalParticlePusher = function (spec) {
  validate(spec);
  const sim = new native.Sim(spec);
  const out = {};
  const rgba = new Uint8Array(4 * spec.nr * spec.nz);
  // out.canvas: an object with the pixels the page would drawImage() (fusionsim.js:154,178)
  Object.defineProperty(out, 'canvas', { get() { sim.render(rgba); return { width: spec.nr, height: spec.nz, data: rgba }; } });
  out.set = function (value) {
    if (value.E) sim.setArray('E', flat3(value.E, spec.nr, spec.nz));
    if (value.B) sim.setArray('B', flat3(value.B, spec.nr, spec.nz));
    if (value.position) sim.setArray('position', flatN(value.position, 3));
    if (value.velocity) sim.setArray('velocity', flatN(value.velocity, 3));
    if (value.sink_mask) sim.setArray('sink_mask', flat2(value.sink_mask));
    if (value.source_pdf) sim.setSourcePdf(flat2(value.source_pdf), value.source_pdf.length, value.source_pdf[0].length);
    if (value.rand) sim.setArray('rand', flatN(value.rand, 4));        // extension: seeding
    if (value.entropy) sim.setArray('entropy', flatN(value.entropy, 4));
  };
  out.addCurrentLoop = (r, z, I) => sim.addCurrentLoop(r, z, I);
  out.addSpindleCuspPlasmaField = (r, B_c, beta_c) => sim.addSpindleCuspPlasmaField(r, B_c, beta_c);
  out.addCurrentZ = (I) => sim.addCurrentZ(I);
  out.addBZ = (Bz) => sim.addBZ(Bz);
  out.addBTheta = (Bt) => sim.addBTheta(Bt);
  out.precalc = () => sim.precalc();
  out.step = () => sim.step();
  out.density = () => sim.density();
  // n iterations of the page loop of fusionsim.js:170-178 (step(); density()) in one call: same results; a scene
  // whose frame is bound by launch latency (the demo scene) replays a captured CUDA graph of 16 frames
  out.runFrames = (n) => sim.runFrames(n);
  // EXTENSION (no reference counterpart): self-consistent electrostatic field solve, include/fusionsim.h
  out.solveFields = function (value) {
    for (const p of ['macro_weight', 'sweeps']) {
      if (typeof value[p] === 'undefined') throw new Error('.' + p + ' <- Non-optional property is undefined!');
      if (typeof value[p] !== 'number') throw new Error('.' + p + ' <- Property does not match any given possible types!');
    }
    sim.solveFields(value.macro_weight, value.sweeps, typeof value.omega === 'number' ? value.omega : 1.0,
                    value.source === 'instant' ? 1 : 0);
  };
  // EXTENSION (no reference counterpart): electromagnetic update on an axisymmetric Yee mesh, include/fusionsim.h.
  // A frame of the loop is halfStep() + density() + emStep(macro_weight): particles and fields advance by the same dt.
  out.halfStep = () => sim.halfStep();
  out.emInit = () => sim.emInit();
  out.emSet = (name, data) => sim.emSet(name, Float64Array.from(data));
  out.emGet = function (name) {
    const extra = { Er: [1, 0], Ez: [0, 1], Bt: [0, 0], Et: [1, 1], Br: [0, 1], Bz: [1, 0] }[name];
    if (!extra) throw new Error('.name <- unknown field ' + name + ' (Er Ez Bt Et Br Bz)');
    const a = new Float64Array((spec.nz + extra[0]) * (spec.nr + extra[1]));
    sim.emGet(name, a);
    return a;
  };
  out.emStep = (macro_weight, with_current) => sim.emStep(macro_weight || 0.0, with_current !== false);
  // checkpoint / restore (extension): everything a run needs to continue bit for bit
  out.checkpoint = function () {
    const n = spec.nparticles * spec.nparticles, nc = spec.nr * spec.nz;
    const get = (name, len) => { const a = new Float64Array(len); sim.getArray(name, a); return a; };
    return { position: get('position', 4 * n), velocity: get('velocity', 3 * n), rand: get('rand', 4 * n),
             E: get('E', 3 * nc), B: get('B', 3 * nc), moments01_avg: get('moments01_avg', 4 * nc) };
  };
  out.restore = function (ck) {  // on a simulation created with the same spec and static tables
    const toIJ = (a) => {  // [cell = i + j*nr][3] as getArray returns it -> value.E[i][j][k] order
      const o = new Float64Array(a.length);
      for (let j = 0; j < spec.nz; j++) for (let i = 0; i < spec.nr; i++) for (let k = 0; k < 3; k++)
        o[(i * spec.nz + j) * 3 + k] = a[(i + j * spec.nr) * 3 + k];
      return o;
    };
    sim.setArray('E', toIJ(ck.E)); sim.setArray('B', toIJ(ck.B));
    sim.precalc();
    sim.setState(ck.position, ck.velocity, ck.rand);
    sim.setField('moments01_avg', ck.moments01_avg);
  };
  // accessors (extension)
  out.getPositions = () => { const a = new Float64Array(4 * spec.nparticles * spec.nparticles); sim.getArray('position', a); return a; };
  out.getVelocities = () => { const a = new Float64Array(3 * spec.nparticles * spec.nparticles); sim.getArray('velocity', a); return a; };
  out.getField = (name, len) => { const a = new Float64Array(len); sim.getArray(name, a); return a; };
  return out;
};
