// ComponentContainer.cc -- see ComponentContainer.H
#include "ComponentContainer.H"

#include "global.H"

namespace exphost {

Component *ComponentContainer::add(std::unique_ptr<Component> c)
{
  Component *p = c.get();
  owned.push_back(std::move(c));
  components.push_back(p);
  return p;
}

void ComponentContainer::add_interaction(Component *from, Component *on)
{
  for (auto &i : interaction)
    if (i->c == from) { i->l.push_back(on); return; }
  auto i = std::make_shared<Interaction>();
  i->c = from;
  i->l.push_back(on);
  interaction.push_back(i);
}

void ComponentContainer::compute_expansion(unsigned mlevel)
{
  for (auto c : components) {
    c->force->set_multistep_level(mlevel);
    c->force->determine_coefficients(c);
  }
}

void ComponentContainer::compute_potential(unsigned mlevel)
{
  // self forces: zero acc/pot/potext of the active levels, then add this component's own force
  for (auto c : components) {
    c->ZeroPotAccel((int)mlevel);
    c->force->set_multistep_level(mlevel);
    c->force->get_acceleration_and_potential(c);
  }
  // interactions: inter->c's basis evaluated on the other component's particles with the
  // coefficients it already holds (SetExternal: no recomputation)
  for (auto &inter : interaction) {
    for (auto other : inter->l) {
      inter->c->force->SetExternal();
      inter->c->force->set_multistep_level(mlevel);
      inter->c->force->get_acceleration_and_potential(other);
      inter->c->force->ClearExternal();
    }
  }
  // external forces are outside the force loop proper.
  // Compute new centre(s) (src/ComponentContainer.cc:952-999)
  int clev = centerlevl < 0 ? (int)multistep / 2 : centerlevl;      // :42-45
  if (clev > (int)multistep) clev = (int)multistep;
  if (mactive.empty() || mactive[mstep][clev]) fix_positions();
}

void ComponentContainer::fix_positions()
{
  mtot = 0.0;
  for (int k = 0; k < 3; k++) gcom[k] = gcov[k] = 0.0;
  for (auto c : components) {
    c->fix_positions();
    mtot += c->mtot;
    for (int k = 0; k < 3; k++) gcom[k] += c->com[k];
    for (int k = 0; k < 3; k++) gcov[k] += c->cov[k];
  }
}

void ComponentContainer::multistep_reset()
{
  for (auto c : components) c->force->multistep_reset();
}

} // namespace exphost
